"""GPU experiment: accuracy of the GEMM modes over the full 1000-step OIL loop (B poses), against the
CUDA-core float32 mode on the same device and (for a 64-pose subset) the numpy oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zedo_oracle as zo
import zedo_release_b200 as zr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
W = zo.make_weights(seed=0)
ds = zo.make_synthetic_dataset(B, seed=1234, n_clusters=1)
cfg = zo.H36M_ZEDO_CFG
plan = zr.ScorePlan(W, n_joints=17, max_batch=B, device=0)
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
uv, K = t(ds["db_2d"][:, :, :2]), t(ds["camera_param"])
x0 = t(zo.init_hypothesis(ds["clusters"], 0, B))
R, T, x_rot, qs = zr.ipo_fit(x0, uv, K, cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"], cfg["IPO_minScaleT"], cfg["IPO_maxScaleT"], 500)
ts = zo.oil_time_grid(1000)[:steps]
gt = ds["db_3d"].astype(np.float64)
res = {}
dumps = [0, 9, 99, 199, 499, steps - 1]
dumps = sorted(set(d for d in dumps if d < steps))
for mode in ("fp32", "split3", "fp8lo", "split2", "fp16"):
    x, Tm = x_rot.clone(), T.clone()
    conf = t(ds["db_2d"][:, :, 2])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = plan.oil_loop(x, Tm, uv, K, conf, ts, phase_switch=steps // 5, dump_steps=dumps, mode=mode)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[mode] = dict(x=x.cpu().numpy(), dump=d.cpu().numpy(), sec=dt)
def mp(x): return np.array([zo.mpjpe(x[n], gt[n]) for n in range(B)])
base = res["fp32"]
out = {"B": B, "steps": steps}
for mode in ("split3", "fp8lo", "split2", "fp16"):
    r = res[mode]
    drift = [float(np.abs(r["dump"][k] - base["dump"][k]).max() / np.abs(base["dump"][k]).max()) for k in range(len(dumps))]
    dm = mp(r["x"]) - mp(base["x"])
    out[mode] = dict(drift_vs_fp32=dict(zip(map(str, dumps), drift)), mpjpe_diff_mm_mean=float(np.abs(dm).mean() * 1e3),
                     mpjpe_diff_mm_max=float(np.abs(dm).max() * 1e3), mpjpe_mean_diff_mm=float(abs(dm.mean()) * 1e3), sec=r["sec"])
out["fp32"] = dict(sec=base["sec"], mpjpe_mean_m=float(mp(base["x"]).mean()))
# oracle on a 64-pose subset, from the same (R, T)
n = 64
xo, To, _ = zo.oil_loop_schedule(W, x_rot.cpu().numpy()[:n], T.cpu().numpy()[:n].reshape(n, 1, 3), ds["db_2d"][:n, :, :2], ds["camera_param"][:n], ds["db_2d"][:n, :, 2].copy(), ts, steps // 5)
for mode in ("fp32", "split3", "fp8lo", "split2", "fp16"):
    xm = res[mode]["x"][:n]
    dmo = np.array([zo.mpjpe(xm[i], gt[i]) - zo.mpjpe(xo[i], gt[i]) for i in range(n)])
    out[mode]["vs_oracle64"] = dict(drift=float(np.abs(xm - xo).max() / np.abs(xo).max()), mpjpe_diff_mm_mean=float(np.abs(dmo).mean() * 1e3), mpjpe_diff_mm_max=float(np.abs(dmo).max() * 1e3),
                                    aggregate_mpjpe_diff_mm=float(abs(dmo.mean()) * 1e3))
out["note"] = ("per-pose |dMPJPE| between two float32 CPU implementations (numpy oracle vs the torch reference) on these same 64 poses: "
               "mean 0.147 mm, max 1.71 mm, drift 1.0e-3, aggregate 0.0007 mm (oracle/gen_golden.py-style run, MPJPE level 2.18 m)")
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"accuracy_modes_B{B}_s{steps}.json"), "w"), indent=1)
