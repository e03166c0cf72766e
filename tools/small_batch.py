"""GPU experiment: OIL-loop time per step at small batches, 64-wide tiles (latency mode) vs 256-wide."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr
W = zo.make_weights(seed=0)
res = {}
for B in (128, 886, 1024, 2048, 4096, 8192, 16384, 32768):
    ds = zo.make_synthetic_dataset(B, seed=1)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
    uv, K, conf = t(ds["db_2d"][:, :, :2]), t(ds["camera_param"]), t(ds["db_2d"][:, :, 2])
    x0 = t(ds["db_3d"] + 0.1); T0 = t(zo.init_translation(ds["db_2d"][:, :, :2], ds["camera_param"], 3.0).reshape(B, 3))
    ts = zo.oil_time_grid()
    for _ in range(2):
        x, T = x0.clone(), T0.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        plan.oil_loop(x, T, uv, K, conf.clone(), ts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[B] = dt * 1e3  # us per step (1000 steps -> ms total == us/step)
    plan.close()
print(os.environ.get("ZEDO_SMALL_TILES", "default"), os.environ.get("ZEDO_GEOM", "auto"), json.dumps(res))
