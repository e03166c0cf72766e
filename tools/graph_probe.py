"""GPU experiment: OIL loop at small batches -- stream launches (7 dependent kernels per step with programmatic
dependent launch) vs one CUDA-graph replay of the whole loop (captured through torch on the same stream)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zedo_release_b200 as zr
from zedo_release_b200 import synthetic as sy, _native as nat
W = sy.make_weights(seed=0)
out = {}
for B in (256, 1024, 2048, 8192):
    ds = sy.make_synthetic_dataset(B, seed=1)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
    uv, K, conf = t(ds["db_2d"][:, :, :2]), t(ds["camera_param"]), t(ds["db_2d"][:, :, 2])
    x0 = t(ds["db_3d"] + 0.1)
    T0 = torch.zeros((B, 3), device="cuda"); T0[:, 2] = 5.0
    ts = zr.linspace_schedule(0.1, 0.01, 1000)
    x, T = x0.clone(), T0.clone()
    res = {}
    for pdl in (1, 0):
        nat.set_option(nat.OPT_PDL, pdl)
        for _ in range(2):
            x.copy_(x0); T.copy_(T0)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            plan.oil_loop(x, T, uv, K, conf, ts)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[f"stream_pdl{pdl}_us_per_step"] = dt * 1e3
        ref = x.clone()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            plan.oil_loop(x, T, uv, K, conf, ts)  # tables built on this stream
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            t0 = time.perf_counter()
            x.copy_(x0); T.copy_(T0)
            with torch.cuda.graph(g, stream=s):
                plan.oil_loop(x, T, uv, K, conf, ts)
            res[f"capture_pdl{pdl}_ms"] = (time.perf_counter() - t0) * 1e3
            for _ in range(2):
                x.copy_(x0); T.copy_(T0)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                g.replay()
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
            res[f"graph_pdl{pdl}_us_per_step"] = dt * 1e3
            res[f"graph_pdl{pdl}_equal"] = bool(torch.equal(x, ref))
            del g
    nat.set_option(nat.OPT_PDL, 1)
    out[B] = res
    plan.close()
print(json.dumps(out))
