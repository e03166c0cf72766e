#!/usr/bin/env python
"""Benchmark of the ZeDO per-pose optimisation loop on B200 (BASELINE.json metric: poses/s of the
full loop = per pose S x (500 IPO iterations + 1000 OIL steps), device-timed).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this implementation, BASELINE configs[1]
    python bench.py --config c1|c2|c3|c4|c5 [--hypo S] ...            # the other BASELINE configs (see CONFIGS)
    python bench.py --impl reference [--steps K] [--warmup W]         # the reference's own PyTorch path on the host CPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...             # one rank per GPU

A "step" is one pass of the hot path over one batch of synthetic input with random-init weights.  Default =
BASELINE configs[1]: H36M format, J=17, hypo=1, 262,144 poses per GPU ("weak": the per-GPU batch is fixed).  c3 / c4 /
c5 fix the TOTAL number of poses and shard it over the ranks ("strong").  Every configuration runs through the
product's sharded runner (zedo_release_b200.parallel.run_sharded): contiguous pose shards (the rule of
lib/dataset/EvaSampler.py:79-112), no collective inside the loop, then MPJPE / argmin on the device and ONE NCCL
all_gather_into_tensor per result tensor (results [N,S,J,3], err_min [N], argmin [N]).

One JSON line is printed by rank 0:
  value     poses/s with the shard's inputs resident in HBM: IPO + OIL loop only, CUDA events, max over ranks
  e2e       the same metric through the public API from PINNED HOST buffers: H2D of the shard's inputs, IPO + OIL,
            evaluation, the NCCL gather of results / errors / argmin (N > 1) and the D2H of the gathered results on
            rank 0 are all inside the timed region; bytes are counted from the tensors moved
  roofline  the dominant kernel (1024x1024 hidden layer) timed live with CUDA events on the launching stream
  cpu_baseline (N = 1)  the UNMODIFIED reference (oracle/_ref, staged by oracle/fetch_ref.py) on the host cores, bounded sample
`--impl reference` prints the reference arm's line: the same reference CPU run as its own arm.
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
# OMP_NUM_THREADS=1 to its workers, and BLAS / torch read it at import -- so it is overridden here, first.
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poses/sec (full diffusion+opt loop: 500 IPO iterations + 1000 OIL steps per pose, device-timed)"
FLOP_PER_POSE_HIDDEN_LAYER = 2 * 1024 * 1024          # one 1024x1024 layer, per pose

# BASELINE.json `configs` -> bench presets.  total = poses of the whole job (strong scaling), per_gpu = weak scaling.
CONFIGS = {
    "c1": dict(dataset="h36m", joints=17, net="score", hypo=1, total=1024,
               label="BASELINE configs[0]: H36M J=17 hypo=1, 1,024 poses"),
    "c2": dict(dataset="h36m", joints=17, net="score", hypo=1, per_gpu=262144,
               label="BASELINE configs[1]: H36M J=17 hypo=1, 262,144 poses per GPU"),
    "c3": dict(dataset="h36m", joints=17, net="score", hypo=50, total=65536,
               label="BASELINE configs[2]: H36M J=17 S=50 cluster-initialised hypotheses, 65,536 poses sharded"),
    "c4": dict(dataset="pw3d", joints=17, net="score", hypo=1, total=1048576,
               label="BASELINE configs[3]: 3DPW-format in-the-wild config (17 key joints, IPO_T=8), 1,048,576 poses sharded"),
    "c5": dict(dataset="syrip", joints=12, net="control", hypo=1, total=65536,
               label="BASELINE configs[4]: infant (SyRIP J=12) fine-tuned (Control) architecture, 65,536 poses sharded"),
}


def flop_per_pose_step(J, control):
    """live per-pose GEMM flops of one network forward."""
    return 2 * (3 * 3 * J * 1024 + 9 * 1024 * 1024) if control else 2 * (2 * 3 * J * 1024 + 4 * 1024 * 1024)


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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, world):
    c = CONFIGS[args.config]
    B = args.poses_local
    return {"workload": f"{c['label']}{'' if args.is_preset else ' [sizes overridden on the command line]'}; "
                        f"{args.dataset} J={args.joints} hypo={args.hypo}, random-init "
                        f"{'control (infant)' if args.net == 'control' else 'concat'} score net; "
                        f"{args.oil_steps} OIL steps + 500 IPO iterations per pose and hypothesis",
            "preset": args.config, "poses_per_gpu": B, "poses_total": args.poses_total, "hypotheses": args.hypo,
            "oil_steps": args.oil_steps, "ipo_iterations": 500, "gemm_mode": args.mode,
            "parallelism": f"pose-sharded x{world} (EvaSampler rule), no collective in the loop; one NCCL "
                           f"all_gather_into_tensor each of results / err_min / argmin at the end",
            "l2": "inputs larger than L2 (GBs of activations per layer pass)" if B * args.hypo >= 65536 else
                  "small batch: the working set fits L2, as in the reference's own runs"}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch path on the host cores (oracle/ref_runner.py over oracle/_ref)
# ---------------------------------------------------------------------------------------------------
def reference_cpu_sample(n_poses=1024, ipo_iters=500, n_phase1=6, n_phase2=18, oil_steps=1000, threads=None):
    """One bounded sample of the reference path on `n_poses` H36M-format poses, hypo = 1: the full 500 IPO iterations
    (run/opt_main.py:180-195) + n_phase1 fixed-T OIL steps + n_phase2 solved-T OIL steps of the 1000-step schedule
    (run/opt_main.py:202-220), extrapolated to 200 + 800 steps (every step of a phase does the same work).
    Returns (poses/s, sample description, detail dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_runner as rr
    import zedo_oracle as zo
    if not rr.available():
        return None, "reference sources not staged (python oracle/fetch_ref.py in the build container)", {}
    R = rr.load()
    torch = R.torch
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    W = zo.make_weights(seed=0)
    ds = zo.make_synthetic_dataset(n_poses, seed=1234, n_clusters=1)
    cfg = dict(zo.H36M_ZEDO_CFG)
    model = rr.build_model(R, W, "cpu")
    sde, sampling_fn = rr.make_sampling_fn(R, "cpu", n_poses)
    cond = torch.tensor(ds["db_2d"][:, :, :2]).float()
    conf = torch.tensor(ds["db_2d"][:, :, 2]).float()
    K = torch.tensor(ds["camera_param"]).float()
    x0 = torch.tensor(zo.init_hypothesis(ds["clusters"], 0, n_poses))
    t0 = time.perf_counter()
    rot, T = rr.ipo(R, x0, cond, K, cfg, "cpu", iters=ipo_iters)
    t_ipo = (time.perf_counter() - t0) * (500 / ipo_iters)
    x = rot.bmm(x0.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
    t0 = time.perf_counter()
    res, T1, _ = rr.oil(R, model, sampling_fn, sde, x.clone(), T, cond, K, conf, "cpu", steps=oil_steps,
                        phase_switch=oil_steps, n_run=n_phase1)
    t_p1 = (time.perf_counter() - t0) / n_phase1
    t0 = time.perf_counter()
    rr.oil(R, model, sampling_fn, sde, torch.tensor(res), T1, cond, K, conf, "cpu", steps=oil_steps, phase_switch=0,
           n_run=n_phase2)
    t_p2 = (time.perf_counter() - t0) / n_phase2
    total = t_ipo + t_p1 * (oil_steps // 5) + t_p2 * (oil_steps - oil_steps // 5)
    sample = (f"UNMODIFIED reference (oracle/_ref: lib/algorithms/advanced/*, driver body run/opt_main.py:166-222) in "
              f"PyTorch {torch.__version__} on {threads} CPU threads, {n_poses} poses hypo=1: {ipo_iters} IPO iterations "
              f"+ {n_phase1} fixed-T and {n_phase2} solved-T OIL steps of the {oil_steps}-step schedule timed, "
              f"extrapolated to {oil_steps // 5} + {oil_steps - oil_steps // 5} steps")
    return n_poses / total, sample, dict(t_ipo_s=t_ipo, ms_per_step_phase1=1e3 * t_p1, ms_per_step_phase2=1e3 * t_p2,
                                         seconds_per_1024_pose_run=total, threads=threads)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 2)):  # thread pools, allocator, first-touch: a short sample is a full warm-up
        reference_cpu_sample(256, ipo_iters=20, n_phase1=2, n_phase2=2)
    vals, sample, info = [], "", {}
    for _ in range(max(1, args.steps)):
        v, sample, info = reference_cpu_sample(1024)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": sample}), flush=True)
            return
        vals.append(v)
    v = float(np.mean(vals))
    B = args.poses_local
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "poses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # one step of the stated workload (B poses x hypo) at the measured rate -- the sample itself is bounded
        "ms_per_step": 1e3 * B * args.hypo / v,
        "higher_is_better": True, "scaling": "weak" if "per_gpu" in CONFIGS[args.config] else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "poses/s", "cores": info.get("threads", cores), "kind": "reference",
                         "sample": sample, "detail": info},
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
    from zedo_release_b200 import parallel
    from zedo_release_b200 import synthetic as sy

    S, J = args.hypo, args.joints
    n_total = args.poses_total_for(world)
    lo, hi = zr.shard_range(n_total, rank, world)
    B = hi - lo
    base_cfg = {"h36m": sy.H36M_ZEDO_CFG, "pw3d": sy.PW3D_ZEDO_CFG, "mini": sy.MINI_ZEDO_CFG,
                "syrip": sy.SYRIP_ZEDO_CFG}[args.dataset]
    cfg = dict(base_cfg)
    cfg["OIL_iterations"] = args.oil_steps
    infant = args.dataset in ("mini", "syrip")
    run_kw = dict(phase_switch=int(0.95 * args.oil_steps), ray_init=True, use_conf=False,
                  pelvis=(0, 3) if args.dataset == "syrip" else (0, 0), root_relative=False,
                  per_hypothesis_cluster=False) if infant else {}
    # every rank generates its own shard (seed offset by rank); clusters are shared
    ds = sy.make_synthetic_dataset(B, n_joints=J, seed=1234 + rank, n_clusters=S)
    if rank != 0:
        ds["clusters"] = sy.make_synthetic_dataset(2, n_joints=J, seed=1234, n_clusters=S)["clusters"]
    control = args.net == "control"
    # hypotheses are stacked along the batch axis when the plan has room (up to ~512k rows = 4.3 GB of activations)
    cap = B * max(1, min(S, 524288 // max(B, 1)))
    plan = zr.ScorePlan(sy.make_weights(seed=0, n_joints=J, control=control), n_joints=J, max_batch=cap, device=local,
                        kind=zr._native.NET_CONTROL if control else zr._native.NET_SCORE_FC_ADV)
    plan.reserve(args.oil_steps, args.mode)
    h_db2d = torch.from_numpy(ds["db_2d"]).pin_memory()
    h_K = torch.from_numpy(ds["camera_param"]).pin_memory()
    h_cl = torch.from_numpy(ds["clusters"]).pin_memory()
    h_gt = torch.from_numpy(ds["db_3d"].astype(np.float64)).pin_memory()
    h_out = torch.empty((n_total if rank == 0 else 1, S, J, 3), dtype=torch.float32).pin_memory()
    h_err = torch.empty((n_total if rank == 0 else 1,), dtype=torch.float64).pin_memory()
    h_idx = torch.empty((n_total if rank == 0 else 1,), dtype=torch.int32).pin_memory()
    d_db2d, d_K, d_cl = h_db2d.to(dev), h_K.to(dev), h_cl.to(dev)

    def step_resident():
        return zr.run_pose_optimisation(plan, d_db2d, d_K, d_cl, cfg, hypo=S, mode=args.mode, b_global=n_total,
                                        **run_kw)

    def step_e2e():
        """The public sharded call from pinned host buffers: H2D, IPO + OIL, evaluation, NCCL gather, D2H on rank 0."""
        res, (err, idx) = parallel.run_sharded(plan, h_db2d, h_K, h_cl, cfg, hypo=S, mode=args.mode, gt=h_gt,
                                               protocol2=True, local_shard=True, n_total=n_total, **run_kw)
        if rank == 0:
            h_out.copy_(res, non_blocking=True)
            h_err.copy_(err, non_blocking=True)
            h_idx.copy_(idx, non_blocking=True)
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
    step_e2e()  # NCCL communicator set-up, pinned staging: outside the timed regions
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
    finite = bool(torch.isfinite(h_out).all()) if rank == 0 else True

    # sharded == unsharded: rank 0 re-runs a slice of its shard on its own (same global batch size for the IPO loss
    # mean) and compares it with the rows that came back through the gather -- bit-exact, rows are independent
    shard_ok = None
    if rank == 0:
        n_chk = min(B, 96)
        sl = zr.run_pose_optimisation(plan, d_db2d[:n_chk].contiguous(), d_K[:n_chk].contiguous(), d_cl, cfg, hypo=S,
                                      mode=args.mode, b_global=n_total, **run_kw)
        torch.cuda.synchronize()
        shard_ok = bool(torch.equal(sl.cpu(), h_out[:n_chk]))

    # K4 / K5 on their own (they are 0.1 % of the step, so they are timed separately: 5 launches each, CUDA events)
    def time_call(fn, reps=5):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    uv_d = d_db2d[:, :, :2].contiguous()
    x0_d = (d_cl - d_cl[:, 0:1])[0].expand(B, J, 3).contiguous()
    ipo_ms = time_call(lambda: zr.ipo_fit(x0_d, uv_d, d_K, cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"],
                                          cfg["IPO_minScaleT"], cfg["IPO_maxScaleT"], 500, b_global=n_total))
    pred_d = x0_d[:, None].contiguous()
    gt_d = h_gt.to(dev)
    eval_ms = time_call(lambda: zr.eval_multi(pred_d, gt_d, protocol2=True))

    poses_total_steps = n_total * args.steps
    value = poses_total_steps / (ms_total / 1e3)
    e2e = poses_total_steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    hid_ms, hid_n = prof["hidden_layer"]
    rows = B * min(S, max(1, cap // max(B, 1)))  # rows one hidden-layer launch carries
    achieved_tflops = (FLOP_PER_POSE_HIDDEN_LAYER * rows) / (hid_ms / 1e3) / 1e12 if hid_ms > 0 else None
    ncu_traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(tp) and rows == 262144:
        try:
            with open(tp) as f:
                summ = json.load(f)
            key = {"split3": "hidden_layer_dram_bytes_per_launch",
                   "fp8lo": "hidden_layer_fp8lo_dram_bytes_per_launch"}.get(args.mode)
            ncu_traffic = summ.get(key) if key else None
            traffic_src = "profiles/ncu_summary.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` " \
                          "capture of this kernel at this size (not measured in this run)"
        except Exception:
            ncu_traffic = None
    gather_bytes = (n_total * S * J * 3 * 4 + n_total * 12) if world > 1 else 0
    h2d = (h_db2d.numel() * 4 + h_K.numel() * 4 + h_cl.numel() * 4 + h_gt.numel() * 8) * world
    d2h = h_out.numel() * 4 + h_err.numel() * 8 + h_idx.numel() * 4 if rank == 0 else 0
    n_prod = {"split3": 3, "fp8lo": 4, "split2": 2, "fp16": 1}.get(args.mode, 0)
    # the HBM-bound kernels of a step against the measured copy bandwidth: ALGORITHMIC bytes per row of a launch
    #  over the live CUDA-event time of this run.  first layer: 64-column operand read (256 B) + the output
    # images a later layer reads (fp8lo: hi16 + hi8 + lo8 + lo16 = 6 B per channel; split3: 4 B); post_dense: hi16 + lo16
    # of 1024 channels read + 64 float32 written; geometry: SURVEY 8(d)'s 672 B of pose state + eps read (256 B) + the
    # first layer's operand written (256 B).
    alg_bytes = {"first_layer": 256 + 1024 * (6 if args.mode == "fp8lo" else 4), "post_dense": 1024 * 4 + 256,
                 "geometry": 672 + 256 + 256}
    other_hbm = {}
    for k, per_row in alg_bytes.items():
        if k in prof and prof[k][0] > 0 and not control:
            gbs = per_row * rows / (prof[k][0] / 1e3) / 1e9
            other_hbm[k] = {"algorithmic_bytes_per_row": per_row, "GBps": gbs, "frac_of_hbm_peak": gbs / peaks["hbm"]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, sample, info = reference_cpu_sample(1024)
            if v is not None:
                cpu = {"value": v, "unit": "poses/s", "cores": info["threads"], "kind": "reference", "sample": sample,
                       "detail": info}
        fps = flop_per_pose_step(J, control)
        line = {
            "metric": METRIC, "value": value, "unit": "poses/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak" if "per_gpu" in CONFIGS[args.config] else "strong", "vs_baseline": None,
            "dtype": {"split3": "f32 (fp16 hi/lo 3-product split on tcgen05, f32 accumulate)",
                      "fp8lo": "f32 (fp16 main product + e4m3 low-order products on tcgen05, f32 accumulate)"
                      }.get(args.mode, args.mode), "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": "poses/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "nccl_gather_bytes_per_step": int(gather_bytes), "ms_per_step": ms_e2e / args.steps,
                    "includes": "H2D of each rank's shard (db_2d, K, clusters, gt), IPO + OIL, eval_multi protocol 2, "
                                "NCCL all_gather_into_tensor of results/err_min/argmin, D2H of the gathered tensors on rank 0"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": f"layer_tc2_kernel<{n_prod},GN_SILU> (1024x1024 hidden layer, "
                                                     "tcgen05 cta_group::2)",
                         "achieved": achieved_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved_tflops / peaks["tflops"]) if achieved_tflops else None,
                         "traffic": ncu_traffic, "traffic_source": traffic_src, "peak_source": peaks["src"],
                         "algorithmic_flop_per_launch": FLOP_PER_POSE_HIDDEN_LAYER * rows,
                         "mma_issue_factor": {"split3": 3, "fp8lo": 2, "split2": 2}.get(args.mode, 1),
                         "avg_launch_ms": hid_ms, "launches_timed": hid_n,
                         "other_kernels_ms": {k: v[0] for k, v in prof.items() if k != "hidden_layer"},
                         "other_kernels_hbm": other_hbm,
                         "other_kernels": {
                             "ipo_fit (K4, 500 Adam iterations in registers)": {
                                 "ms": ipo_ms, "bound": "fp32 ALU (serial per pose)", "poses": B,
                                 "hbm_GBps_algorithmic": 424 * B / (ipo_ms / 1e3) / 1e9, "hbm_frac": 424 * B / (ipo_ms / 1e3) / 1e9 / peaks["hbm"]},
                             "eval_multi protocol 2 (K5, fp64 Procrustes)": {
                                 "ms": eval_ms, "bound": "fp64 ALU (3x3 Jacobi SVD per pose and hypothesis)", "poses": B,
                                 "hbm_GBps_algorithmic": (2 * J * 12 + 12) * B / (eval_ms / 1e3) / 1e9,
                                 "hbm_frac": (2 * J * 12 + 12) * B / (eval_ms / 1e3) / 1e9 / peaks["hbm"]}}},
            "oil_pose_steps_per_s": n_total * args.oil_steps * S / (ms_total / args.steps / 1e3),
            "loop_tflops_algorithmic": fps * n_total * args.oil_steps * S / (ms_total / args.steps / 1e3) / 1e12,
            "results_finite": finite,
            "sharded_equals_unsharded_slice": shard_ok,
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs preset")
    ap.add_argument("--poses", type=int, default=None, help="override: poses per GPU (weak scaling)")
    ap.add_argument("--poses-total", type=int, default=None, help="override: poses of the whole job (strong scaling)")
    ap.add_argument("--hypo", type=int, default=None)
    ap.add_argument("--oil-steps", type=int, default=1000, help="OIL steps per pose (reference: 1000)")
    ap.add_argument("--mode", default="fp8lo", choices=["split3", "fp8lo", "split2", "fp16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dataset", default=None, choices=["h36m", "pw3d", "mini", "syrip"],
                    help="ZeDO config block (IPO key joints / axes / scale clamp; infant configs switch phase at 95%%)")
    ap.add_argument("--joints", type=int, default=None)
    ap.add_argument("--net", default=None, choices=["score", "control"])
    args = ap.parse_args()
    preset = CONFIGS[args.config]
    args.is_preset = all(getattr(args, k) is None for k in ("poses", "poses_total", "hypo", "dataset", "joints", "net"))
    for k in ("dataset", "joints", "net", "hypo"):
        if getattr(args, k) is None:
            setattr(args, k, preset[k])
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def poses_total_for(w):
        if args.poses is not None:
            return args.poses * w
        if args.poses_total is not None:
            return args.poses_total
        return preset["per_gpu"] * w if "per_gpu" in preset else preset["total"]

    args.poses_total_for = poses_total_for
    args.poses_total = poses_total_for(world)
    args.poses_local = -(-args.poses_total // world)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
