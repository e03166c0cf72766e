"""GPU experiment: teacher-forced error of the step components against the numpy oracle (1024 poses)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr
B = 1024
W = zo.make_weights(seed=0)
ds = zo.make_synthetic_dataset(B, seed=1234)
t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
uv, K, conf = ds["db_2d"][:, :, :2], ds["camera_param"], ds["db_2d"][:, :, 2].copy()
np.clip(conf, 1e-4, 1, out=conf)
x = (ds["db_3d"] * 1.3 + 0.05).astype(np.float32)
T_in = zo.init_translation(uv, K, 3.0)
plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
def rel(a, b): return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
def rms(a, b): return float(np.sqrt(np.mean((np.asarray(a, np.float64) - b) ** 2)) / np.sqrt(np.mean(np.asarray(b, np.float64) ** 2)))
out = {}
# reference solution of the LS problem in float64 (the "truth" both float32 solvers approximate)
g_o, T_o = zo.gradient_field(uv, x, K, conf=conf.copy())
g_g, T_g = zr.grad_field(t(uv), t(x), t(K), conf=t(conf))
x64, uv64, K64 = x.astype(np.float64), uv.astype(np.float64), K.astype(np.float64)
Kinv = np.linalg.inv(K64); h = np.concatenate([uv64, np.ones((B, 17, 1))], -1)
ray = np.einsum("bij,bnj->bni", Kinv, h); ray /= ray[:, :, 2:]
w = (conf.astype(np.float64) ** 2)[:, :, None]
A = np.zeros((B, 34, 3)); b = np.zeros((B, 34, 1))
b[:, 0::2] = (x64[:, :, 0:1] - x64[:, :, 2:3] * ray[:, :, 0:1]) * w; b[:, 1::2] = (x64[:, :, 1:2] - x64[:, :, 2:3] * ray[:, :, 1:2]) * w
A[:, 0::2, 0] = -w[:, :, 0]; A[:, 0::2, 2] = ray[:, :, 0] * w[:, :, 0]; A[:, 1::2, 1] = -w[:, :, 0]; A[:, 1::2, 2] = ray[:, :, 1] * w[:, :, 0]
T64 = np.linalg.solve(A.transpose(0, 2, 1) @ A, A.transpose(0, 2, 1) @ b).transpose(0, 2, 1)
T64[T64[:, :, 2] < 0] *= -1
out["T_solve"] = dict(gpu_vs_oracle=rel(T_g.cpu().numpy(), T_o), gpu_vs_f64=rel(T_g.cpu().numpy(), T64), oracle_vs_f64=rel(T_o, T64),
                      rms_gpu_vs_f64=rms(T_g.cpu().numpy(), T64), rms_oracle_vs_f64=rms(T_o, T64))
out["grad_solveT"] = dict(gpu_vs_oracle=rel(g_g.cpu().numpy(), g_o))
g_o2, _ = zo.gradient_field(uv, x, K, t=T_in, conf=conf.copy())
g_g2, _ = zr.grad_field(t(uv), t(x), t(K), conf=t(conf), T=t(T_in))
out["grad_fixedT"] = dict(gpu_vs_oracle=rel(g_g2.cpu().numpy(), g_o2), rms=rms(g_g2.cpu().numpy(), g_o2))
for tt in (0.1, 0.05, 0.012):
    _, xm_o = zo.pc_sampler_step(W, x, np.float32(tt))
    for mode in ("fp32", "split3"):
        _, xm_g = plan.sde_step(t(x), float(np.float32(tt)), mode=mode)
        out[f"sde_step_t{tt}_{mode}"] = dict(rel=rel(xm_g.cpu().numpy(), xm_o), rms=rms(xm_g.cpu().numpy(), xm_o),
                                            rel_of_increment=float(np.abs(xm_g.cpu().numpy() - xm_o).max() / np.abs(xm_o - x).max()))
    e_o = zo.score_forward(W, x, np.float32(tt) * np.float32(999))
    for mode in ("fp32", "split3"):
        out[f"net_t{tt}_{mode}"] = dict(rel=rel(plan.forward(t(x), float(np.float32(tt) * np.float32(999)), mode=mode).cpu().numpy(), e_o))
print(json.dumps(out, indent=1))
