"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck): every kernel once or twice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr
from zedo_release_b200 import _native as nat
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
ds = zo.make_synthetic_dataset(B, seed=2, n_clusters=2)
cfg = dict(zo.H36M_ZEDO_CFG); cfg["IPO_iterations"] = 5
for kind, W in ((nat.NET_SCORE_FC_ADV, zo.make_weights(0)), (nat.NET_CONTROL, zo.make_weights(3, control=True))):
    plan = zr.ScorePlan(W, n_joints=17, max_batch=2 * B, device=0, kind=kind)
    for mode in ("split3", "fp16", "fp32"):
        out = plan.forward(t(ds["db_3d"]), 42.0, mode=mode)
    res = zr.run_pose_optimisation(plan, t(ds["db_2d"]), t(ds["camera_param"]), t(ds["clusters"]), cfg, hypo=2, steps=3)
    e, i = zr.eval_multi(res, t(ds["db_3d"], torch.float64), protocol2=True)
    zr.pck_auc(res, t(ds["db_3d"], torch.float64), select=i)
    zr.hypothesis_std(res)
    torch.cuda.synchronize()
    plan.close()
big = zr.ScorePlan(zo.make_weights(0), n_joints=17, max_batch=4096, device=0)  # CTA-pair path (more than 18 row tiles)
x = t(np.random.default_rng(0).normal(0, 0.3, (4096, 17, 3)).astype(np.float32))
big.forward(x, 10.0)
# the loop on precomputed rays (large-batch geometry form) and as one CUDA graph, on a side stream, ragged batch
ds2 = zo.make_synthetic_dataset(4001, seed=4)
uv, K, conf = t(ds2["db_2d"][:, :, :2]), t(ds2["camera_param"]), t(ds2["db_2d"][:, :, 2])
xl, Tl = t(ds2["db_3d"]), t(zo.init_translation(ds2["db_2d"][:, :, :2], ds2["camera_param"], 3.0).reshape(4001, 3))
nat.set_option(nat.OPT_GEOM_KERNEL, 3)
big.oil_loop(xl, Tl, uv, K, conf, zo.oil_time_grid()[:6], phase_switch=2, dump_steps=(1, 5))
side = torch.cuda.Stream()
torch.cuda.synchronize()
nat.set_option(nat.OPT_GRAPH, 1)
with torch.cuda.stream(side):
    for _ in range(2):
        big.oil_loop(xl, Tl, uv, K, conf, zo.oil_time_grid()[:6], phase_switch=2)
    side.synchronize()
nat.set_option(nat.OPT_GRAPH, 0)
nat.set_option(nat.OPT_GEOM_KERNEL, 0)
torch.cuda.synchronize()
big.close()
print("sanitize smoke done, finite:", bool(torch.isfinite(res).all()))
