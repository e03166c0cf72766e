"""GPU experiment: OIL loop at small batches -- stream launches (7 dependent kernels per step with programmatic
dependent launch) vs the product's graph mode (zedo_set_option(ZEDO_OPT_GRAPH, 1): the whole loop call captured once,
then one cudaGraphLaunch per loop).  Reports us per step of the direct loop, the cost of the first (capturing) call
and us per step of a replay, and that the results are bit-identical."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zedo_release_b200 as zr
from zedo_release_b200 import synthetic as sy, _native as nat
W = sy.make_weights(seed=0)
out = {}
s = torch.cuda.Stream()
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
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        def run():
            x.copy_(x0); T.copy_(T0)
            s.synchronize(); t0 = time.perf_counter()
            plan.oil_loop(x, T, uv, K, conf, ts)
            s.synchronize()
            return time.perf_counter() - t0
        for pdl in (1, 0):
            nat.set_option(nat.OPT_PDL, pdl)
            nat.set_option(nat.OPT_GRAPH, 0)
            run()
            res[f"stream_pdl{pdl}_us_per_step"] = min(run(), run()) * 1e3
            ref = x.clone()
            nat.set_option(nat.OPT_GRAPH, 1)
            res[f"graph_pdl{pdl}_first_call_ms"] = run() * 1e3  # capture + instantiate + the loop itself
            res[f"graph_pdl{pdl}_us_per_step"] = min(run(), run()) * 1e3
            res[f"graph_pdl{pdl}_equal"] = bool(torch.equal(x, ref))
    nat.set_option(nat.OPT_PDL, 1)
    nat.set_option(nat.OPT_GRAPH, 0)
    out[B] = res
    plan.close()
print(json.dumps(out))
