"""CPU study (numpy emulation, no GPU): how much of the loop's final-pose deviation each tensor-core operand
format would cause, to decide which GEMM modes are worth building.

Every mode runs the oracle's OIL loop (oracle/zedo_oracle.py, `forward=` hook) with the four 1024x1024 layers
replaced by an emulation of the operand rounding; products of 16-/8-bit operands are exact in float32 and the
accumulation is numpy's float32 sgemm, like the float32 TMEM accumulator.  The truth is the same loop with float64
GEMMs on the unrounded float32 operands.  Modes:
  f32      numpy float32 sgemm on the full operands (what the oracle and the reference do)
  split3   A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, fp16 hi/lo pairs            (3 fp16 passes; the shipped parity mode)
  split2   A_hi.(W_hi + W_lo)                                             (2 fp16 passes)
  fp8lo    A_hi.W_hi in fp16 + e4m3(A_lo 2^11).e4m3(W_hi 2^-11) + e4m3(A_hi).e4m3(W_lo)
           (1 fp16 pass + 2 fp8 passes = 2 fp16-pass equivalents at the 2x fp8 rate)
  fp16     A_hi.W_hi                                                      (1 pass)
Usage: python tools/precision_study.py [poses] [steps] [damp] > profiles/r01_precision_study.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import zedo_oracle as zo

f32 = np.float32


def fp16(x):
    return x.astype(np.float16).astype(f32)


def e4m3(x):
    """round to nearest e4m3 (4 significant bits, min normal 2^-6, subnormal step 2^-9, max 448)"""
    x = np.asarray(x, dtype=f32)
    ax = np.abs(x)
    m, e = np.frexp(ax)  # ax = m 2^e, m in [0.5, 1)
    q = np.ldexp(np.round(m * 16.0) / 16.0, e)
    sub = np.round(ax / 2.0 ** -9) * 2.0 ** -9
    out = np.where(ax < 2.0 ** -6, sub, q)
    return (np.sign(x) * np.minimum(out, 448.0)).astype(f32)


class Layer:
    def __init__(self, w):
        s = 2.0 ** np.floor(np.log2(384.0 / np.abs(w).max()))  # max|w| s in [256, 512) like the plan's packing
        while np.abs(w).max() * s >= 512:
            s /= 2
        while np.abs(w).max() * s < 256:
            s *= 2
        self.s = f32(s)
        ws = (w * self.s).astype(f32)
        self.w = w
        self.w64 = w.astype(np.float64)
        self.hi = fp16(ws)
        self.lo = fp16(ws - self.hi)
        # the fp8lo images are packed 2^6 higher (pack_weight_f8: max |w| s in [2^14, 2^15)), which keeps e4m3(W_hi 2^-11)
        # out of the e4m3 subnormals down to max / 2^10
        self.s8 = f32(64.0)
        ws8 = (ws * self.s8).astype(f32)
        hi_s = fp16(ws8)
        self.hi8 = e4m3(hi_s * f32(2.0 ** -11))
        self.lo8 = e4m3(fp16(ws8 - hi_s))
        self.hilo = (self.hi + self.lo).astype(f32)  # exact in float32 (22 bits)

    def apply(self, a, mode):
        if mode == "f64":
            return (a.astype(np.float64) @ self.w64.T).astype(f32)
        if mode == "f32":
            return a @ self.w.T
        a_hi = fp16(a)
        if mode == "fp16":
            return (a_hi @ self.hi.T) / self.s
        if mode == "split2":
            return (a_hi @ self.hi.T + a_hi @ self.lo.T) / self.s
        a_lo = fp16(a - a_hi)
        if mode == "split3":
            return (a_hi @ self.hi.T + a_lo @ self.hi.T + a_hi @ self.lo.T) / self.s
        if mode == "fp8lo":
            return (a_hi @ self.hi.T + (e4m3(a_lo * f32(2.0 ** 11)) @ self.hi8.T + e4m3(a_hi) @ self.lo8.T) / self.s8) / self.s
        raise ValueError(mode)


def make_forward(W, mode):
    names = [f"b{k}_dense{j}" for k in (1, 2) for j in (1, 2)]
    layers = {n: Layer(W[n + ".weight"]) for n in names}

    def forward(W, x, t999, n_blocks=2):
        Bx = x.shape[0]
        temb = zo.time_embed(W, np.asarray(t999, dtype=f32))
        tp = lambda n: zo.linear(temb, W[n + ".weight"], W[n + ".bias"])
        h = zo.linear(x.reshape(Bx, -1).astype(f32), W["pre_dense.weight"], W["pre_dense.bias"]) + tp("pre_dense_t")
        h = zo.silu(zo.group_norm(h, W["pre_gnorm.weight"], W["pre_gnorm.bias"]))
        for k in (1, 2):
            h1 = layers[f"b{k}_dense1"].apply(h, mode) + W[f"b{k}_dense1.bias"] + tp(f"b{k}_dense1_t")
            h1 = zo.silu(zo.group_norm(h1.astype(f32), W[f"b{k}_gnorm1.weight"], W[f"b{k}_gnorm1.bias"]))
            h2 = layers[f"b{k}_dense2"].apply(h1, mode) + W[f"b{k}_dense2.bias"] + tp(f"b{k}_dense2_t")
            h2 = zo.silu(zo.group_norm(h2.astype(f32), W[f"b{k}_gnorm2.weight"], W[f"b{k}_gnorm2.bias"]))
            h = (h + h2).astype(f32)
        return zo.linear(h, W["post_dense.weight"], W["post_dense.bias"]).reshape(x.shape).astype(f32)

    return forward


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    DAMP = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    W = zo.make_weights(seed=0)
    if DAMP != 1.0:  # realistic-scale network output (tests/golden/oil_small.npz uses 0.05)
        W["post_dense.weight"] = (W["post_dense.weight"] * f32(DAMP)).astype(f32)
        W["post_dense.bias"] = (W["post_dense.bias"] * f32(DAMP)).astype(f32)
    ds = zo.make_synthetic_dataset(B, seed=1234, n_clusters=1)
    uv, K = ds["db_2d"][:, :, :2], ds["camera_param"]
    x0 = zo.init_hypothesis(ds["clusters"], 0, B)
    cfg = zo.H36M_ZEDO_CFG
    R, T = zo.ipo_fit(x0, uv, K, cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"], cfg["IPO_minScaleT"],
                      cfg["IPO_maxScaleT"], 100)
    x_rot = np.einsum("bij,bkj->bki", R, x0).astype(f32)
    ts = zo.oil_time_grid()[:STEPS] if STEPS < 1000 else zo.oil_time_grid()
    res, out = {}, {}
    x_in = ds["db_3d"][:4] + f32(0.1)
    for mode in ("f64", "f32", "split3", "fp8lo", "split2", "fp16"):
        t0 = time.time()
        fwd = make_forward(W, mode)
        x, Tt, _ = zo.oil_loop_schedule(W, x_rot, T, uv, K, ds["db_2d"][:, :, 2].copy(), ts, len(ts) // 5, forward=fwd)
        res[mode] = x
        print(mode, f"{time.time() - t0:.0f}s", file=sys.stderr, flush=True)
    gt = ds["db_3d"]
    mp = lambda r: np.linalg.norm(r - gt, axis=-1).mean(axis=1)  # per pose, metres
    truth = res["f64"]
    for mode in ("f32", "split3", "fp8lo", "split2", "fp16"):
        d = np.abs(mp(res[mode]) - mp(truth)) * 1e3
        fwd_err = float(np.abs(make_forward(W, mode)(W, x_in, f32(49.95)) - make_forward(W, "f64")(W, x_in, f32(49.95))).max()
                        / np.abs(make_forward(W, "f64")(W, x_in, f32(49.95))).max())
        out[mode] = {"forward_rel_err_vs_f64": fwd_err,
                     "final_pose_rel_drift": float(np.abs(res[mode] - truth).max() / np.abs(truth).max()),
                     "per_pose_dMPJPE_mm": {"mean": float(d.mean()), "p99": float(np.quantile(d, 0.99)), "max": float(d.max())},
                     "aggregate_dMPJPE_mm": float(abs(mp(res[mode]).mean() - mp(truth).mean()) * 1e3)}
    print(json.dumps({"what": "numpy emulation of tensor-core operand formats in the OIL loop; deviations of the final "
                              "poses from the float64-GEMM run of the same float32 loop", "poses": B, "steps": len(ts),
                      "post_dense_damping": DAMP, "mpjpe_level_m": float(mp(truth).mean()), "modes": out}, indent=1))


if __name__ == "__main__":
    main()
