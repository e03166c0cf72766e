"""Summarise where the issue slots of a kernel go, from `ncu --set full --import-source on` captures.

Usage: python tools/ncu_stalls.py OUT.md REPORT.ncu-rep [REPORT.ncu-rep ...]
For every captured launch: duration, clock, issue-slot utilisation, pipe utilisation, DRAM bytes, the warp-stall
samples by reason and by opcode, the executed-instruction mix, and the instructions that collect the most samples
(SASS; the library is built with -lineinfo).  The .ncu-rep files are scratch; the markdown is the tracked evidence."""
import collections
import csv
import subprocess
import sys


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


RAW = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
       ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
       ("smsp__inst_executed.sum", "warp instructions"),
       ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
       ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
       ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
       ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
       ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
       ("launch__registers_per_thread", "registers/thread"),
       ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
       ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes")]


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    md = ["# Issue-slot / stall analysis of the loop kernels (ncu --set full, SASS-level sampling)\n",
          "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <n> -c <m> python bench.py "
          "--steps 1 --warmup 0 --oil-steps 4 --no-cpu` (262,144 poses), summarised by `tools/ncu_stalls.py`.  "
          "Times under ncu are cold-cache, serialised and at whatever clock the idle-then-busy GPU picks "
          "(see the SM clock column); the in-loop figures are in `r01_bench_n1.json`.\n"]
    for rep in reps:
        raw = page(rep, "raw")
        hdr, units = raw[0], raw[1]
        src = page(rep, "source")
        kernels, cur = [], None
        for r in src:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "rows": []}
                kernels.append(cur)
            elif r and r[0] == "Address":
                cur["hdr"] = r
            elif cur is not None and len(r) > 10:
                cur["rows"].append(r)
        for li, r in enumerate(raw[2:]):
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            md.append(f"\n## `{d['Kernel Name'][:110]}`\n")
            md.append("| metric | value |\n|---|---|")
            for key, label in RAW:
                if key in d and d[key] not in ("", "n/a"):
                    md.append(f"| {label} | {d[key]} {u[key]} |")
            stalls = [(h.split("issue_stalled_")[1].split("_per_issue")[0], float(d[h])) for h in hdr
                      if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h] not in ("", "n/a")]
            stalls.sort(key=lambda kv: -kv[1])
            md.append("\nWarps stalled per issued instruction, by reason: " +
                      ", ".join(f"{k} {v:.2f}" for k, v in stalls if v >= 0.05) + "\n")
            def norm(n):
                return (n.replace("zedo::", "").replace("(int)", "").replace("(bool)", "").replace("void ", "")
                        .replace(" ", ""))
            match = [k for k in kernels if norm(k["name"]) == norm(d["Kernel Name"])]
            per = len(kernels) // max(1, len(raw) - 2)  # source blocks per captured launch (ncu 2025: two views each)
            if per >= 1 and norm(kernels[li * per]["name"]) == norm(d["Kernel Name"]):
                match = [kernels[li * per]]  # launch order
            elif not match or li > 0:
                continue  # a source page that only carries the first launch of a multi-launch report
            k = match[0]
            ix = {n: i for i, n in enumerate(k["hdr"])}
            by_s, by_i = collections.Counter(), collections.Counter()
            for row in k["rows"]:
                toks = [t for t in row[ix["Source"]].split() if not t.startswith("@")]
                op = toks[0].split(".")[0]
                by_s[op] += int(row[ix["# Samples"]])
                by_i[op] += int(row[ix["Instructions Executed"]])
            tot_s, tot_i = sum(by_s.values()), sum(by_i.values())
            md.append("Executed warp instructions by opcode (millions): " +
                      ", ".join(f"{o} {n / 1e6:.1f}" for o, n in by_i.most_common(14)) + f" — total {tot_i / 1e6:.1f}\n")
            md.append("Stall samples by opcode (% of samples): " +
                      ", ".join(f"{o} {100 * n / tot_s:.1f}" for o, n in by_s.most_common(10)) + "\n")
            cols = [c for c in k["hdr"] if c.startswith("stall_") and "Not Issued" not in c]
            top = sorted(k["rows"], key=lambda row: -int(row[ix["# Samples"]]))[:12]
            md.append("| hottest instructions (SASS) | samples % | executed | main stall reasons |\n|---|---|---|---|")
            for row in top:
                st = sorted(((c[6:], int(row[ix[c]])) for c in cols if int(row[ix[c]]) > 0), key=lambda kv: -kv[1])[:3]
                md.append(f"| `{row[ix['Source']].strip()[:70]}` | {100 * int(row[ix['# Samples']]) / tot_s:.1f} | "
                          f"{row[ix['Instructions Executed']]} | {', '.join(f'{a} {b}' for a, b in st)} |")
    open(out_path, "w").write("\n".join(md) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
