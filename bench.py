#!/usr/bin/env python
"""Benchmark of the ZeDO per-pose optimisation loop on B200 (BASELINE.json metric: poses/s of the
full loop = per pose S x (500 IPO iterations + 1000 OIL steps), device-timed).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # this implementation
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU port of the reference path
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU

A "step" is one pass of the hot path over one batch of synthetic H36M-format input: BASELINE
config 1 of `configs` index 1 -- J=17, hypo=1, 262,144 poses per GPU -- with random-init weights.
Poses are independent, so N GPUs run N shards with no data-path collective ("weak" scaling: the
per-GPU batch is fixed); the only exchange is one gather of the results, inside the e2e timing.

One JSON line is printed by rank 0 (keys: see the task contract): `value` = poses/s with inputs
resident in HBM, CUDA-event timed, max over ranks; `e2e` = the same loop through the public API
with pinned-host inputs/outputs copied inside the timed region; `roofline` = the dominant kernel
(hidden 1024x1024 layer, tcgen05 3-product split) timed live with CUDA events on the launching
stream inside the timed region; `cpu_baseline` = the numpy oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU legs (the `--impl reference` arm, `cpu_baseline` at N = 1) use every host core.  torchrun exports
# OMP_NUM_THREADS=1 to its workers, and BLAS reads it when numpy is imported -- so it is overridden here, first.
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poses/sec (full diffusion+opt loop: 500 IPO iterations + 1000 OIL steps per pose, device-timed)"
FLOP_PER_POSE_HIDDEN_LAYER = 2 * 1024 * 1024          # one 1024x1024 layer, per pose
FLOP_PER_POSE_STEP = 2 * (51 * 1024 + 4 * 1024 * 1024 + 1024 * 51)  # 8,597,504


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d["bf16_tflops_sustained"]), hbm=float(d["hbm_gbs"]), src="measured (sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle port of the reference path (oracle/ is only ever used here as the baseline)
# ---------------------------------------------------------------------------------------------------
def _cpu_port_shard(job):
    """One worker of the CPU arm: the numpy oracle port on poses [lo, hi) with one BLAS thread (the workers
    together use every core; rows are independent, the IPO loss mean uses the global batch size)."""
    lo, hi, n_poses, oil_steps_sampled, ipo_iters, oil_steps_total = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zedo_oracle as zo
    W = zo.make_weights(seed=0)
    ds = zo.make_synthetic_dataset(n_poses, seed=1234, n_clusters=1)
    cfg = zo.H36M_ZEDO_CFG
    uv, K = ds["db_2d"][lo:hi, :, :2], ds["camera_param"][lo:hi]
    x0 = zo.init_hypothesis(ds["clusters"], 0, hi - lo)
    t0 = time.perf_counter()
    R, T = zo.ipo_fit(x0, uv, K, cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"], cfg["IPO_minScaleT"],
                      cfg["IPO_maxScaleT"], iters=ipo_iters, b_global=n_poses)
    t_ipo = time.perf_counter() - t0
    x = np.einsum("bij,bnj->bni", R, x0).astype(np.float32)
    ts = zo.oil_time_grid(oil_steps_total)[:oil_steps_sampled]
    t0 = time.perf_counter()
    zo.oil_loop_schedule(W, x, T, uv, K, ds["db_2d"][lo:hi, :, 2].copy(), ts, oil_steps_total // 5)
    return t_ipo, time.perf_counter() - t0


def cpu_port_poses_per_s(n_poses=1024, oil_steps_sampled=100, ipo_iters=500, oil_steps_total=1000, workers=None):
    """poses/s of the numpy port of the reference path on the host cores: the batch is split over one process
    per core (the phases run concurrently, so each phase costs its slowest worker)."""
    import multiprocessing as mp
    workers = max(1, min(workers or (os.cpu_count() or 1), n_poses))
    bounds = [n_poses * i // workers for i in range(workers + 1)]
    jobs = [(bounds[i], bounds[i + 1], n_poses, oil_steps_sampled, ipo_iters, oil_steps_total) for i in range(workers)]
    if workers == 1:
        parts = [_cpu_port_shard(jobs[0])]
    else:
        # spawn, not fork: the GPU arm calls this with a live CUDA context and helper threads
        with mp.get_context("spawn").Pool(workers) as pool:
            parts = pool.map(_cpu_port_shard, jobs)
    t_ipo, t_oil = max(p[0] for p in parts), max(p[1] for p in parts)
    total = t_ipo + t_oil * (oil_steps_total / oil_steps_sampled)
    sample = (f"{n_poses} poses (BASELINE config 0 shape) split over {workers} worker processes: {ipo_iters} IPO "
              f"iterations + {oil_steps_sampled} of {oil_steps_total} OIL steps, OIL time scaled "
              f"x{oil_steps_total / oil_steps_sampled:g}")
    return n_poses / total, sample, dict(t_ipo_s=t_ipo, t_oil_sampled_s=t_oil, workers=workers)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(max(0, args.warmup - 2)):  # numpy/BLAS warm-up; each full sample costs ~10 s
        cpu_port_poses_per_s(256, 5, 20)
    vals, sample = [], ""
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        v, sample, info = cpu_port_poses_per_s(1024, 50)
        cores = info["workers"]
        vals.append(v)
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "poses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "H36M J=17 hypo=1, random-init concat score net (BASELINE configs[1] shape), "
                               "CPU port of the reference path on a bounded sample", "poses_per_gpu": args.poses,
                   "oil_steps": 1000, "ipo_iterations": 500},
        "cpu_baseline": {"value": v, "unit": "poses/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        entry.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    import zedo_release_b200 as zr
    from zedo_release_b200 import synthetic as sy

    B, S, J = args.poses, args.hypo, args.joints
    base_cfg = {"h36m": sy.H36M_ZEDO_CFG, "pw3d": sy.PW3D_ZEDO_CFG, "mini": sy.MINI_ZEDO_CFG,
                "syrip": sy.SYRIP_ZEDO_CFG}[args.dataset]
    cfg = dict(base_cfg)
    cfg["OIL_iterations"] = args.oil_steps
    infant = args.dataset in ("mini", "syrip")
    run_kw = dict(phase_switch=int(0.95 * args.oil_steps), ray_init=True, use_conf=False,
                  pelvis=(0, 3) if args.dataset == "syrip" else (0, 0)) if infant else {}
    ds = sy.make_synthetic_dataset(B, n_joints=J, seed=1234 + rank, n_clusters=S)
    control = args.net == "control"
    # hypotheses are stacked along the batch axis when the plan has room (up to ~512k rows = 4.3 GB of activations)
    cap = B * max(1, min(S, 524288 // B))
    plan = zr.ScorePlan(sy.make_weights(seed=0, n_joints=J, control=control), n_joints=J, max_batch=cap, device=local,
                        kind=zr._native.NET_CONTROL if control else zr._native.NET_SCORE_FC_ADV)
    h_db2d = torch.from_numpy(ds["db_2d"]).pin_memory()
    h_K = torch.from_numpy(ds["camera_param"]).pin_memory()
    h_cl = torch.from_numpy(ds["clusters"]).pin_memory()
    h_out = torch.empty((B, S, J, 3), dtype=torch.float32).pin_memory()
    d_db2d, d_K, d_cl = h_db2d.to(dev), h_K.to(dev), h_cl.to(dev)
    b_global = B * world  # the IPO loss is a mean over the whole (global) batch (run/opt_main.py:191)

    def step_resident():
        return zr.run_pose_optimisation(plan, d_db2d, d_K, d_cl, cfg, hypo=S, mode=args.mode, b_global=b_global,
                                        **run_kw)

    def step_e2e():
        a = h_db2d.to(dev, non_blocking=True)
        k = h_K.to(dev, non_blocking=True)
        c = h_cl.to(dev, non_blocking=True)
        res = zr.run_pose_optimisation(plan, a, k, c, cfg, hypo=S, mode=args.mode, b_global=b_global, **run_kw)
        h_out.copy_(res, non_blocking=True)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = zr._native.launch_count()
    plan.profile(True, stride=53)
    ms_total = timed(step_resident, args.steps)
    prof = plan.profile_read()
    plan.profile(False)
    launches = zr._native.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    if world > 1:  # the one exchange of the path: gather of the results (here: per-shard checksum)
        chk = torch.tensor([float(h_out.double().abs().mean())], device=dev, dtype=torch.float64)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
    finite = bool(torch.isfinite(h_out).all())

    poses_total = B * world * args.steps
    value = poses_total / (ms_total / 1e3)
    e2e = poses_total / (ms_e2e / 1e3)
    peaks = load_peaks()
    hid_ms, hid_n = prof["hidden_layer"]
    achieved_tflops = (FLOP_PER_POSE_HIDDEN_LAYER * B) / (hid_ms / 1e3) / 1e12 if hid_ms > 0 else None
    ncu_traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                summ = json.load(f)
            # the captured figure belongs to the default mode's kernel at the default size; fp8lo has its own
            key = {"split3": "hidden_layer_dram_bytes_per_launch",
                   "fp8lo": "hidden_layer_fp8lo_dram_bytes_per_launch"}.get(args.mode)
            ncu_traffic = summ.get(key) if (key and B == 262144) else None
        except Exception:
            ncu_traffic = None
    h2d = h_db2d.numel() * 4 + h_K.numel() * 4 + h_cl.numel() * 4
    d2h = h_out.numel() * 4

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, sample, info = cpu_port_poses_per_s(1024, 50)
            cpu = {"value": v, "unit": "poses/s", "cores": info["workers"], "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": "poses/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"split3": "f32 (fp16 hi/lo 3-product split on tcgen05, f32 accumulate)",
                                                         "fp8lo": "f32 (fp16 main product + e4m3 low-order products on tcgen05, f32 accumulate)"
                                                         }.get(args.mode, args.mode), "data": "synthetic",
            "config": {"workload": f"{args.dataset} J={J} hypo={S}, {B} synthetic poses per GPU, random-init "
                                   f"{'control (infant)' if control else 'concat'} score net"
                                   f"{' (BASELINE configs[1])' if (args.dataset, J, S, B, control) == ('h36m', 17, 1, 262144, False) else ''}"
                                   f"; {args.oil_steps} OIL steps + 500 IPO iterations per pose",
                       "poses_per_gpu": B, "hypotheses": S, "oil_steps": args.oil_steps, "ipo_iterations": 500,
                       "gemm_mode": args.mode, "parallelism": f"pose-sharded x{world}, no data-path collective",
                       "l2": "inputs larger than L2 (2.1 GB of activations per layer pass)"},
            "e2e": {"value": e2e, "unit": "poses/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": f"layer_tc2_kernel<{ {'split3': 3, 'fp8lo': 4, 'split2': 2, 'fp16': 1}.get(args.mode, 0) },GN_SILU> "
                                                     "(1024x1024 hidden layer, tcgen05 cta_group::2)",
                         "achieved": achieved_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved_tflops / peaks["tflops"]) if achieved_tflops else None,
                         "traffic": ncu_traffic, "peak_source": peaks["src"],
                         "algorithmic_flop_per_launch": FLOP_PER_POSE_HIDDEN_LAYER * B,
                         "mma_issue_factor": {"split3": 3, "fp8lo": 2, "split2": 2}.get(args.mode, 1),
                         "avg_launch_ms": hid_ms, "launches_timed": hid_n,
                         "other_kernels_ms": {k: v[0] for k, v in prof.items() if k != "hidden_layer"}},
            "oil_pose_steps_per_s": B * world * args.oil_steps * S / (ms_total / args.steps / 1e3),
            "loop_tflops_algorithmic": (2 * (3 * 3 * J * 1024 + 9 * 1024 * 1024) if control else
                                        2 * (2 * 3 * J * 1024 + 4 * 1024 * 1024)) * B * world * args.oil_steps * S
            / (ms_total / args.steps / 1e3) / 1e12,
            "results_finite": finite,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=262144, help="poses per GPU (BASELINE configs[1])")
    ap.add_argument("--hypo", type=int, default=1)
    ap.add_argument("--oil-steps", type=int, default=1000, help="OIL steps per pose (reference: 1000)")
    ap.add_argument("--mode", default="split3", choices=["split3", "fp8lo", "split2", "fp16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dataset", default="h36m", choices=["h36m", "pw3d", "mini", "syrip"],
                    help="ZeDO config block (IPO key joints / axes / scale clamp; infant configs switch phase at 95%%)")
    ap.add_argument("--joints", type=int, default=17)
    ap.add_argument("--net", default="score", choices=["score", "control"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
