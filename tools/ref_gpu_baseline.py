"""The reference's own eager-PyTorch path (oracle/_ref, unmodified) timed on the GPU and on the host CPU:
BASELINE.md section 5 rows "reference PyTorch eager on B200" and "reference on CPU".  fp32, TF32 off (torch default),
per-step host round trip as shipped (sampling.py:515,525).  Writes one JSON object to stdout.

    python tools/ref_gpu_baseline.py [--cpu-full]   # --cpu-full: also the full C1 run on the host cores (~1 min)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import ref_runner as rr
import zedo_oracle as zo


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    R = rr.load()
    W = zo.make_weights(seed=0)
    cfg = dict(zo.H36M_ZEDO_CFG)
    out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(),
           "tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
    m = rr.build_model(R, W, "cuda")
    # C1: 1,024 poses, the full pipeline (500 IPO + 1000 OIL), twice (first = warm-up)
    ds = zo.make_synthetic_dataset(1024, seed=1234, n_clusters=1)
    for rep in range(2):
        t0 = time.perf_counter()
        res, info = rr.run_pipeline(R, m, ds["db_2d"], ds["camera_param"], ds["clusters"], cfg, "cuda")
        wall = time.perf_counter() - t0
    out["c1_gpu"] = {"poses": 1024, "wall_s": wall, "t_ipo_s": info["t_ipo"], "t_oil_s": info["t_oil"],
                     "poses_per_s": 1024 / wall, "oil_us_per_step": 1e3 * info["t_oil"]}
    # C2 batch: 262,144 poses; 500 IPO iterations are batch-size independent in cost structure -> time 50 of them and
    # 20 OIL steps (10 per phase) and extrapolate linearly (every step does the same work)
    B = 262144
    ds2 = zo.make_synthetic_dataset(B, seed=1234, n_clusters=1)
    t0 = time.perf_counter()
    res, info = rr.run_pipeline(R, m, ds2["db_2d"], ds2["camera_param"], ds2["clusters"], cfg, "cuda", ipo_iters=50,
                                n_run=20, phase_switch=10)
    out["c2_gpu_sample"] = {"poses": B, "t_ipo_50_iters_s": info["t_ipo"], "t_oil_20_steps_s": info["t_oil"]}
    t_full = info["t_ipo"] * 10 + info["t_oil"] * 50
    out["c2_gpu_extrapolated"] = {"seconds_per_262144_poses": t_full, "poses_per_s": B / t_full,
                                  "note": "10 x (50 IPO iterations) + 50 x (10 phase-1 + 10 phase-2 OIL steps)"}
    if "--cpu-full" in sys.argv:
        torch.set_num_threads(os.cpu_count() or 1)
        mc = rr.build_model(R, W, "cpu")
        t0 = time.perf_counter()
        res, info = rr.run_pipeline(R, mc, ds["db_2d"], ds["camera_param"], ds["clusters"], cfg, "cpu")
        wall = time.perf_counter() - t0
        out["c1_cpu_full"] = {"poses": 1024, "threads": torch.get_num_threads(), "wall_s": wall, "t_ipo_s": info["t_ipo"],
                              "t_oil_s": info["t_oil"], "poses_per_s": 1024 / wall}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
