"""GPU report for BASELINE configs[0] size (1,024 poses): final MPJPE of the 1000-step loop against the reference's
own run (tests/golden/c1.npz, written by oracle/gen_golden.py) -- aggregate and per-pose, by GEMM mode, and for the
whole pipeline with this library's IPO.  Usage: python tools/c1_report.py > profiles/rNN_c1_parity.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr
g = np.load(os.path.join(ROOT, "tests", "golden", "c1.npz"))
N = 1024
ds = zo.make_synthetic_dataset(N, seed=int(g["seed"]), n_clusters=1)
W = zo.make_weights(seed=0)
W["post_dense.weight"] = (W["post_dense.weight"] * g["post_scale"]).astype(np.float32)
W["post_dense.bias"] = (W["post_dense.bias"] * g["post_scale"]).astype(np.float32)
dev = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
plan = zr.ScorePlan(W, n_joints=17, max_batch=N, device=0)
uv, K, conf = ds["db_2d"][:, :, :2], ds["camera_param"], ds["db_2d"][:, :, 2]
x0 = zo.init_hypothesis(ds["clusters"], 0, N)
gt = dev(ds["db_3d"].astype(np.float64))
ref = g["mpjpe"]
out = {"poses": N, "oil_steps": 1000, "reference_mpjpe_m": float(ref.mean()),
       "reference_noise_floor": "numpy oracle vs the torch reference on the same inputs: aggregate 0.0005 mm, per-pose "
                                "mean 0.144 mm, max 1.484 mm (tests/golden/PINNING.txt)"}
for mode in ("split3", "fp32", "fp8lo", "split2", "fp16"):
    x, T = dev(np.einsum("bij,bnj->bni", g["R"], x0).astype(np.float32)), dev(g["T"].reshape(N, 3))
    plan.oil_loop(x, T, dev(uv), dev(K), dev(conf), zo.oil_time_grid(), mode=mode)
    err, _ = zr.eval_multi(x[:, None].contiguous(), gt)
    m = err.cpu().numpy()
    d = np.abs(m - ref)
    out[mode] = {"aggregate_dMPJPE_mm": float(abs(m.mean() - ref.mean()) * 1e3), "per_pose_mean_mm": float(d.mean() * 1e3),
                 "per_pose_max_mm": float(d.max() * 1e3)}
res = zr.run_pose_optimisation(plan, dev(ds["db_2d"]), dev(K), dev(ds["clusters"]), zo.H36M_ZEDO_CFG, hypo=1)
for p2 in (False, True):
    e, _ = zr.eval_multi(res, gt, protocol2=p2)
    key = "pa_mpjpe" if p2 else "mpjpe"
    out[f"full_pipeline_own_ipo_{key}_m"] = float(e.mean())
out["full_pipeline_aggregate_dMPJPE_mm"] = float(abs(out["full_pipeline_own_ipo_mpjpe_m"] - ref.mean()) * 1e3)
plan.close()
print(json.dumps(out, indent=1))
