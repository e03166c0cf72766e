"""Turn gpurun_out/launches_rNN.csv + gpurun_out/prof_layer_rNN.ncu-rep into tracked summaries under
profiles/ (the .ncu-rep itself is scratch).  Usage: python tools/ncu_summarise.py r01"""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
summary = {"round": tag}
# ---- launch list --------------------------------------------------------------------------------
lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows, hdr = [], None
    for r in csv.reader(open(lp)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
        elif hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                rows.append((d["Kernel Name"], float(d["Metric Value"].replace(",", ""))))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
    shares = {}
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:90]}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / n / 1e6:.4f} |")
        shares[k[:90]] = dict(launches=n, total_ms=t / 1e6, share=t / tot, avg_ms=t / n / 1e6)
    summary["launch_list"] = shares
    with open(os.path.join(out_dir, f"{tag}_launch_list.md"), "w") as f:
        f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                "`python bench.py --steps 1 --warmup 0 --oil-steps 4 --no-cpu` (262,144 poses; IPO + 4 OIL steps).\n"
                "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n" + "\n".join(lines) + "\n")
# ---- full capture of the layer kernel --------------------------------------------------------------
rp = os.path.join(ROOT, "gpurun_out", f"prof_layer_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
    caps = []
    def num(s, u):
        v = float(s.replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u.strip())
        return v * scale if scale else v
    for r in rows[2:]:
        d = {}
        for w in want:
            if w in idx:
                try:
                    d[w] = num(r[idx[w]], units[idx[w]]) if w != "Kernel Name" else r[idx[w]]
                except ValueError:
                    d[w] = r[idx[w]]
        caps.append(d)
    summary["layer_kernel_captures"] = caps
    hid = [c for c in caps if "layer_tc2_kernel" in c.get("Kernel Name", "")]
    if not hid:
        hid = [c for c in caps if "256, 3, 0" in c.get("Kernel Name", "")]
    # keep the figures of the mode that was not captured this time (bench.py reads one key per GEMM mode)
    prev_path = os.path.join(out_dir, "ncu_summary.json")
    if os.path.exists(prev_path):
        prev = json.load(open(prev_path))
        for k in ("hidden_layer_dram_bytes_per_launch", "hidden_layer_tensor_pipe_active_pct",
                  "hidden_layer_fp8lo_dram_bytes_per_launch", "hidden_layer_fp8lo_tensor_pipe_active_pct",
                  "hidden_layer_fp8lo_l2_to_sm_bytes_per_launch"):
            if k in prev:
                summary[k] = prev[k]
    for tagk, sel in (("", "layer_tc2_kernel<3,"), ("fp8lo_", "layer_tc2_kernel<4,")):
        h = [c for c in hid if sel in c.get("Kernel Name", "")]
        if h:
            summary[f"hidden_layer_{tagk}dram_bytes_per_launch"] = sum(c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"] for c in h) / len(h)
            summary[f"hidden_layer_{tagk}tensor_pipe_active_pct"] = sum(c["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for c in h) / len(h)
            if "l1tex__m_xbar2l1tex_read_bytes.sum" in h[0]:
                summary[f"hidden_layer_{tagk}l2_to_sm_bytes_per_launch"] = sum(c["l1tex__m_xbar2l1tex_read_bytes.sum"] for c in h) / len(h)
    with open(os.path.join(out_dir, f"{tag}_layer_kernel_ncu.md"), "w") as f:
        f.write(f"# ncu --set full, layer_tc_kernel ({tag})\n\n`ncu --set full --clock-control none --import-source on -k regex:layer_tc` "
                "on 262,144 poses.  One row per captured launch.\n\n")
        keys = [w for w in want if w in idx]
        f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for c in caps:
            f.write("| " + " | ".join(f"{c.get(k):.4g}" if isinstance(c.get(k), float) else str(c.get(k)) for k in keys) + " |\n")
json.dump(summary, open(os.path.join(out_dir, f"ncu_summary_{tag}.json"), "w"), indent=1)
json.dump(summary, open(os.path.join(out_dir, "ncu_summary.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k not in ("launch_list", "layer_kernel_captures")}, indent=1))
