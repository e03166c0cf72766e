"""Pin the oracle against the real reference and write the golden vectors.

Runs ONLY in the build container (needs ``/root/reference``; that tree does not exist
on the GPU box).  It imports the reference's own PyTorch modules on CPU, feeds them the
synthetic inputs of ``oracle/zedo_oracle.py``, asserts that every oracle function agrees
with the reference, and stores small input/output fixtures under ``tests/golden/``:

    python oracle/gen_golden.py            # check + (re)write fixtures
    python oracle/gen_golden.py --check    # check only

Shims needed to import the reference here: an empty
``matplotlib``/``matplotlib.pyplot`` (imported unused at simple_zeroshot_opt.py:3), a
stub ``prettytable`` (h36m.py / pw3d.py) and an attribute namespace instead of
``ml_collections.ConfigDict``.
"""
from __future__ import annotations

import argparse
import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("ZEDO_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import zedo_oracle as zo  # noqa: E402


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree {REF} not found: gen_golden.py only runs in the build container")
    for name in ("matplotlib", "matplotlib.pyplot", "prettytable", "h5py"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["prettytable"].PrettyTable = type("PrettyTable", (), {
        "__init__": lambda self, *a, **k: None, "add_row": lambda self, *a, **k: None,
        "field_names": None, "__str__": lambda self: ""})
    sys.path.insert(0, REF)
    import torch
    from lib.algorithms.advanced import sde_lib, sampling, utils as mutils
    from lib.algorithms.advanced.model import ScoreModelFC_Adv
    from lib.algorithms.advanced.control_model import Control_ScoreModelFC_Adv
    from lib.algorithms.advanced import simple_zeroshot_opt as szo
    from lib.utils import transforms
    from lib.dataset.h36m import H36MDataset3D
    from lib.dataset.pw3d import PW3D
    return types.SimpleNamespace(torch=torch, sde_lib=sde_lib, sampling=sampling, mutils=mutils,
                                 ScoreModelFC_Adv=ScoreModelFC_Adv, Control=Control_ScoreModelFC_Adv,
                                 szo=szo, transforms=transforms, H36M=H36MDataset3D, PW3D=PW3D)


def ns(**kw):
    return types.SimpleNamespace(**kw)


def ref_config():
    return ns(
        training=ns(sde="subvpsde", continuous=True, cond_pose_mask_prob=0.0, cond_part_mask_prob=0.0,
                    cond_joint_mask_prob=0.0),
        sampling=ns(method="pc", predictor="euler_maruyama", corrector="none", snr=0.16, n_steps_each=1,
                    probability_flow=True, noise_removal=True),
        model=ns(embedding_type="positional", scale_by_sigma=False, sigma_max=50, sigma_min=0.01,
                 num_scales=1000, beta_min=0.1, beta_max=20.0, t=0.1, ema_rate=0.9999),
        device="cpu",
    )


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


class Checker:
    def __init__(self):
        self.rows = []

    def check(self, name, got, want, tol):
        e = rel_err(got, want)
        ok = e <= tol
        self.rows.append((name, e, tol, ok))
        print(f"  [{'ok' if ok else 'FAIL'}] {name:58s} rel_err={e:.3e} (tol {tol:.0e})")
        return ok

    def bound(self, name, value, tol):
        ok = value <= tol
        self.rows.append((name, float(value), tol, ok))
        print(f"  [{'ok' if ok else 'FAIL'}] {name:58s} value  ={value:.3e} (tol {tol:.0e})")
        return ok

    def all_ok(self):
        return all(r[3] for r in self.rows)


def load_into(torch, module, W):
    sd = {k: torch.tensor(v) for k, v in W.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert set(missing) <= {"sigmas"}, missing
    assert not unexpected, unexpected
    module.eval()
    return module


def pin_formats(R, ck, out):
    # ---- 9. dataset file formats: the reference's own loaders on synthetic files ---------------------------
    print("dataset formats (H36M pickle items, 3DPW npz) through the reference's loaders")
    import pickle
    import tempfile
    dsf = zo.make_synthetic_dataset(30, seed=31, n_clusters=2)
    items = zo.h36m_items_from_arrays(dsf)
    det = {"test": {"joint3d_image": np.concatenate([dsf["db_2d"][:, :, :2] + 1.5, np.zeros((30, 17, 1), np.float32)], -1),
                    "confidence": dsf["db_2d"][:, :, 2:3].copy()}}
    fm = {}
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(io.StringIO()):
        with open(os.path.join(tmp, "h36m_test.pkl"), "wb") as f:
            pickle.dump(items, f)
        with open(os.path.join(tmp, "h36m_sh_dt_ft.pkl"), "wb") as f:
            pickle.dump(det, f)
        np.savez(os.path.join(tmp, "pw3d_test.npz"), **zo.pw3d_npz_from_arrays(dsf))
        h_gt = R.H36M(tmp, "test", gt2d=True, abs_coord=True)
        h_dt = R.H36M(tmp, "test", gt2d=False, abs_coord=True)
        pw_ds = R.PW3D(tmp, "test", abs_coord=True)
        rngf = np.random.default_rng(3)
        preds_f = (dsf["db_3d"][:, None] + rngf.normal(0, 0.03, (30, 4, 17, 3))).astype(np.float32)
        fm["h36m_eval_p1"] = np.float64(h_gt.eval_multi(preds_f, protocol2=False))
        fm["h36m_eval_p2"] = np.float64(h_gt.eval_multi(preds_f, protocol2=True))
        fm["pw3d_eval_p1"] = np.float64(pw_ds.eval_multi(preds_f, protocol2=False))
    ck.check("H36M loader db_3d == absolute synthetic poses (metres)", h_gt.db_3d,
             (dsf["db_3d"].astype(np.float64) + dsf["root"][:, None, :]).astype(np.float32), 1e-6)
    ck.check("H36M loader camera_param", h_gt.camera_param, dsf["camera_param"], 0.0)
    ck.check("PW3D loader db_3d (order_change inverted by the generator)", pw_ds.db_3d, h_gt.db_3d, 1e-6)
    agg_f, _, _ = zo.eval_multi(preds_f.astype(np.float64), (dsf["db_3d"] - dsf["db_3d"][:, 0:1]).astype(np.float64),
                                actions=dsf["actions"])
    ck.check("H36M.eval_multi on the loaded items (action-wise)", agg_f, fm["h36m_eval_p1"], 1e-6)
    fm.update(h36m_db2d_gt=np.asarray(h_gt.db_2d), h36m_db2d_dt=np.asarray(h_dt.db_2d), h36m_db3d=h_gt.db_3d,
              h36m_K=h_gt.camera_param, pw3d_db2d=pw_ds.db_2d, pw3d_db3d=pw_ds.db_3d, pw3d_K=pw_ds.camera_param,
              preds=preds_f, det_xy=det["test"]["joint3d_image"], det_conf=det["test"]["confidence"],
              seed=np.int64(31))
    out["formats"] = fm



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="do not write fixtures")
    ap.add_argument("--formats-only", action="store_true",
                    help="run only the dataset-format section and write formats.npz (PINNING.txt is left alone)")
    ap.add_argument("--only", default="", help="comma-separated fixture names to (re)write; default: all")
    args = ap.parse_args()
    R = _import_reference()
    torch = R.torch
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ck = Checker()
    out = {}
    if args.formats_only:
        pin_formats(R, ck, out)
        if not ck.all_ok():
            raise SystemExit("oracle does NOT match the reference")
        if not args.check:
            np.savez_compressed(os.path.join(GOLD, "formats.npz"), **out["formats"])
            print("wrote tests/golden/formats.npz")
        return

    # ---- 1. SDE scalars and the time grid ---------------------------------------------------
    print("sde / time grid")
    sde = R.sde_lib.subVPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
    ts = torch.linspace(sde.T, 0.01, 1000)
    ck.check("oil_time_grid == torch.linspace", zo.oil_time_grid(), ts.numpy(), 0.0)
    probe_t = torch.tensor([0.1, 0.05, 0.01, float(ts[137]), float(ts[803])])
    x1 = torch.ones(5, 1, 1)
    drift, diff = sde.sde(x1, probe_t)
    _, std = sde.marginal_prob(x1, probe_t)
    beta_o, g_o = zo.subvp_sde_scalars(probe_t.numpy())
    ck.check("subVPSDE.sde diffusion", g_o, diff.numpy(), 2e-5)
    ck.check("subVPSDE.sde drift coefficient", -0.5 * beta_o, drift.numpy().ravel(), 2e-7)
    ck.check("subVPSDE.marginal_prob std (1-exp cancellation: 1 ulp of exp = 3e-5)", zo.subvp_marginal_std(probe_t.numpy()), std.numpy(), 1e-4)
    out["sde"] = dict(t=probe_t.numpy(), diffusion=diff.numpy(), half_beta=-drift.numpy().ravel(),
                      std=std.numpy(), time_grid=ts.numpy())

    # ---- 2. the only known-answer of the reference: the __main__ demo -------------------------
    print("simple_zeroshot_opt __main__ demo")
    key2d = np.array([[[100, 100], [120, 120], [140, 140], [90, 100]]], np.float32)
    key3d0 = np.array([[[1, 1, 3], [1.2, 1.2, 3], [1.4, 1.4, 3], [0.9, 100, 3]]], np.float32)
    Kd = np.array([[[1000, 0, 100], [0, 1000, 100], [0, 0, 1]]], np.float32)
    k3 = torch.FloatTensor(key3d0.copy())
    demo_ref = []
    for i in range(10):
        g = R.szo.gradient_field_gen(torch.FloatTensor(key2d), k3, torch.FloatTensor(Kd))
        demo_ref.append(float(torch.mean(torch.norm(g, dim=-1))))
        k3 += g
    k3o = key3d0.copy()
    demo_or = []
    for i in range(10):
        g, _ = zo.gradient_field(key2d, k3o, Kd)
        demo_or.append(float(np.mean(np.linalg.norm(g, axis=-1))))
        k3o = k3o + g
    assert demo_ref[0] == 53.63671875, demo_ref[0]
    ck.check("demo iteration 0 (53.63671875)", demo_or[0], demo_ref[0], 1e-6)
    ck.check("demo final key3d", k3o, k3.numpy(), 1e-6)
    out["demo"] = dict(key2d=key2d, key3d=key3d0, K=Kd, first_norm=np.float64(demo_ref[0]),
                       final_key3d=k3.numpy())

    # ---- 3. gradient_field_gen, both branches ---------------------------------------------------
    print("gradient_field_gen")
    ds = zo.make_synthetic_dataset(16, seed=7, n_clusters=3)
    uvc = ds["db_2d"].copy()
    uvc[0, 3, 2] = 1.7      # exercises conf > 1 clamp
    uvc[1, 5, 2] = 1e-6     # exercises conf < 1e-4 clamp
    x3 = (ds["db_3d"] + np.random.default_rng(3).normal(0, 0.05, ds["db_3d"].shape)).astype(np.float32)
    Kc = ds["camera_param"].copy()
    Kc[2, 0, 1] = 3.5       # skewed intrinsics -> general 3x3 inverse
    T_in = zo.init_translation(uvc[:, :, :2], Kc, 3)
    cases = {}
    for tag, use_t, use_conf in (("fixedT_conf", True, True), ("solveT_conf", False, True),
                                 ("solveT_noconf", False, False), ("fixedT_noconf", True, False)):
        conf_t = torch.tensor(uvc[:, :, 2].copy()) if use_conf else None
        res = R.szo.gradient_field_gen(torch.tensor(uvc[:, :, :2]), torch.tensor(x3), torch.tensor(Kc),
                                       t=torch.tensor(T_in) if use_t else None, conf=conf_t, returnT=True)
        conf_o = uvc[:, :, 2].copy() if use_conf else None
        g_o, T_o = zo.gradient_field(uvc[:, :, :2], x3, Kc, t=T_in if use_t else None, conf=conf_o)
        ck.check(f"gradient_field_gen {tag}: gradient", g_o, res[0].numpy(), 2e-5)
        ck.check(f"gradient_field_gen {tag}: T", T_o, res[1].numpy(), 2e-5)
        if use_conf:
            ck.check(f"gradient_field_gen {tag}: conf clamped in place", conf_o, conf_t.numpy(), 0.0)
        cases[tag] = dict(g=res[0].numpy(), T=res[1].numpy())
    # a translation with negative z to exercise the sign flip (:93)
    x3_neg = (x3 - np.array([0, 0, 12.0], np.float32)).astype(np.float32)
    res = R.szo.gradient_field_gen(torch.tensor(uvc[:, :, :2]), torch.tensor(x3_neg), torch.tensor(Kc),
                                   returnT=True)
    g_o, T_o = zo.gradient_field(uvc[:, :, :2], x3_neg, Kc)
    ck.check("gradient_field_gen sign-flip case: gradient", g_o, res[0].numpy(), 1e-4)
    ck.check("gradient_field_gen sign-flip case: T", T_o, res[1].numpy(), 2e-5)
    cases["flip"] = dict(g=res[0].numpy(), T=res[1].numpy())
    out["geom"] = dict(db_2d=uvc, x=x3, x_neg=x3_neg, K=Kc, T_in=T_in,
                       **{f"{k}_{n}": v for k, d in cases.items() for n, v in d.items()})

    # ---- 4. score network forward ----------------------------------------------------------------
    print("ScoreModelFC_Adv.forward")
    cfg = ref_config()
    W = zo.make_weights(seed=0)
    model = load_into(torch, R.ScoreModelFC_Adv(cfg, n_joints=17, joint_dim=3, hidden_dim=1024,
                                                embed_dim=512, cond_dim=3), W)
    xb = (ds["db_3d"][:8] + 0.3 * np.random.default_rng(5).normal(0, 1, (8, 17, 3))).astype(np.float32)
    net = {}
    for t in (0.1, 0.05, 0.01):
        lab = torch.ones(8) * torch.tensor(t) * 999
        with torch.no_grad():
            ref = model(torch.tensor(xb), lab, torch.zeros(8, 17, 2), None).numpy()
        got = zo.score_forward(W, xb, np.float32(t) * np.float32(999))
        ck.check(f"score forward t={t}", got, ref, 2e-5)
        net[f"out_{t}"] = ref
    out["net"] = dict(x=xb, weights_seed=np.int64(0), **net)

    # 'fourier' time embedding (configs/default_pose_gen_configs.py:71; model.py:246-250): same weights + gauss_proj.W.
    # The Fourier features multiply log(t) by N(0, 30^2) frequencies, so one ulp of log(t) moves the features by 2e-4;
    # zo.log_f32 is the correctly rounded float32 logarithm, which is what torch.log returns on the whole time grid.
    cfg_f = ref_config()
    cfg_f.model.embedding_type = "fourier"
    Wf = zo.make_weights(seed=0, fourier=True)
    model_f = load_into(torch, R.ScoreModelFC_Adv(cfg_f, n_joints=17, joint_dim=3, hidden_dim=1024,
                                                  embed_dim=512, cond_dim=3), Wf)
    netf = {}
    for t in (0.1, 0.05, 0.01):
        lab = torch.ones(8) * torch.tensor(t) * 999
        with torch.no_grad():
            ref = model_f(torch.tensor(xb), lab, torch.zeros(8, 17, 2), None).numpy()
            emb_ref = model_f.gauss_proj(torch.log(lab[:1])).numpy()
        t999 = np.float32(t) * np.float32(999)
        ck.check(f"fourier embedding t={t}", zo.gaussian_fourier_projection(zo.log_f32(t999), Wf["gauss_proj.W"]),
                 emb_ref, 2e-6)
        ck.check(f"score forward (fourier embedding) t={t}", zo.score_forward(Wf, xb, t999), ref, 2e-5)
        netf[f"out_{t}"] = ref
        netf[f"emb_{t}"] = emb_ref
    grid = zo.oil_time_grid() * np.float32(999)
    ck.check("log_f32 == torch.log on the OIL time grid (bit-exact)", zo.log_f32(grid), torch.log(torch.tensor(grid)).numpy(),
             0.0)
    out["net_fourier"] = dict(x=xb, weights_seed=np.int64(0), **netf)

    # J=12 variant (SyRIP) and the control network (opt_main_infant.py:122-148)
    W12 = zo.make_weights(seed=2, n_joints=12)
    m12 = load_into(torch, R.ScoreModelFC_Adv(cfg, n_joints=12, joint_dim=3, hidden_dim=1024,
                                              embed_dim=512, cond_dim=3), W12)
    x12 = np.random.default_rng(6).normal(0, 0.3, (8, 12, 3)).astype(np.float32)
    with torch.no_grad():
        ref12 = m12(torch.tensor(x12), torch.ones(8) * 49.95, None, None).numpy()
    ck.check("score forward J=12", zo.score_forward(W12, x12, np.float32(49.95)), ref12, 2e-5)
    out["net12"] = dict(x=x12, weights_seed=np.int64(2), t999=np.float32(49.95), out=ref12)

    Wc = zo.make_weights(seed=3, control=True)
    mc = R.Control(cfg, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    mc = load_into(torch, mc, Wc)
    with torch.no_grad():
        refc = mc(torch.tensor(xb), torch.ones(8) * 49.95).numpy()
    ck.check("Control_ScoreModelFC_Adv forward", zo.control_score_forward(Wc, xb, np.float32(49.95)), refc, 2e-5)
    out["control"] = dict(x=xb, weights_seed=np.int64(3), t999=np.float32(49.95), out=refc)

    # ---- 5. one pc_sampler call --------------------------------------------------------------------
    print("pc_sampler")
    sampling_fn = R.sampling.get_sampling_fn(cfg, sde, (8, 17, 3), lambda x: x, 0.01, device="cpu")
    t_i = ts[412]
    trajs, results = sampling_fn(model, condition=torch.zeros(8, 17, 2), gradient=None,
                                 denoise_x=torch.tensor(xb), t=t_i, t_step=412, args=None)
    tr_o, res_o = zo.pc_sampler_step(W, xb, np.float32(t_i))
    ck.check("pc_sampler results", res_o, results, 2e-6)
    ck.check("pc_sampler trajs", tr_o, trajs, 2e-6)
    out["sampler"] = dict(x=xb, t=np.float32(t_i), results=results, trajs=trajs)

    # noise-bearing Euler-Maruyama (probability_flow=False) with an injected z
    score_fn = R.mutils.get_score_fn(sde, model, train=False, continuous=True)
    pred = R.sampling.EulerMaruyamaPredictor(sde, score_fn, probability_flow=False)
    z = np.random.default_rng(8).normal(0, 1, xb.shape).astype(np.float32)
    _orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: torch.tensor(z)
    try:
        with torch.no_grad():
            xr, xmr = pred.update_fn(torch.tensor(xb), torch.ones(8) * t_i, None, None)
            rd = R.sampling.ReverseDiffusionPredictor(sde, score_fn, probability_flow=False)
            xr2, xmr2 = rd.update_fn(torch.tensor(xb), torch.ones(8) * t_i, None, None)
    finally:
        torch.randn_like = _orig
    xo, xmo = zo.euler_maruyama_update(W, xb, np.float32(t_i), z=z, probability_flow=False)
    ck.check("EulerMaruyama (noise) x", xo, xr.numpy(), 2e-6)
    ck.check("EulerMaruyama (noise) x_mean", xmo, xmr.numpy(), 2e-6)
    xo2, xmo2 = zo.reverse_diffusion_update(W, xb, np.float32(t_i), z=z, probability_flow=False)
    ck.check("ReverseDiffusion (noise) x", xo2, xr2.numpy(), 2e-6)
    ck.check("ReverseDiffusion (noise) x_mean", xmo2, xmr2.numpy(), 2e-6)
    out["noise"] = dict(z=z, em_x=xr.numpy(), em_mean=xmr.numpy(), rd_x=xr2.numpy(), rd_mean=xmr2.numpy())

    # ---- 5b. the remaining registered updates with injected noise (sampling.py:208-324): VP / VE SDEs ------------
    print("ancestral / Langevin / ALD with injected noise (VPSDE, VESDE)")
    vp = R.sde_lib.VPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
    ve = R.sde_lib.VESDE(sigma_min=0.01, sigma_max=50.0, N=1000, T=0.1)
    sf_vp = R.mutils.get_score_fn(vp, model, train=False, continuous=True)
    sf_ve = R.mutils.get_score_fn(ve, model, train=False, continuous=True)
    zs = [np.random.default_rng(20 + k).normal(0, 1, xb.shape).astype(np.float32) for k in range(2)]

    def with_noise(fn):
        it = iter(zs)
        _o = torch.randn_like
        torch.randn_like = lambda x, *a, **k: torch.tensor(next(it))
        try:
            with torch.no_grad():
                return tuple(v.numpy() for v in fn())
        finally:
            torch.randn_like = _o

    vt = torch.ones(8) * t_i
    nz = {}
    nz["anc_vp_x"], nz["anc_vp_mean"] = with_noise(lambda: R.sampling.AncestralSamplingPredictor(vp, sf_vp).update_fn(
        torch.tensor(xb), vt, None, None))
    nz["anc_ve_x"], nz["anc_ve_mean"] = with_noise(lambda: R.sampling.AncestralSamplingPredictor(ve, sf_ve).update_fn(
        torch.tensor(xb), vt, None, None))
    nz["lang_x"], nz["lang_mean"] = with_noise(lambda: R.sampling.LangevinCorrector(vp, sf_vp, 0.16, 2).update_fn(
        torch.tensor(xb), vt, None, None))
    nz["ald_x"], nz["ald_mean"] = with_noise(lambda: R.sampling.AnnealedLangevinDynamics(vp, sf_vp, 0.16, 1).update_fn(
        torch.tensor(xb), vt, None, None))
    tf = np.float32(t_i)
    for tag, got in (("anc_vp", zo.ancestral_update_vp(W, xb, tf, zs[0])), ("anc_ve", zo.ancestral_update_ve(W, xb, tf, zs[0])),
                     ("lang", zo.langevin_update_vp(W, xb, tf, zs, snr=0.16, n_steps=2)),
                     ("ald", zo.langevin_update_vp(W, xb, tf, zs, snr=0.16, n_steps=1, ald=True))):
        ck.check(f"{tag} (injected noise) x", got[0], nz[f"{tag}_x"], 5e-6)
        ck.check(f"{tag} (injected noise) x_mean", got[1], nz[f"{tag}_mean"], 5e-6)
    try:
        R.sampling.LangevinCorrector(sde, score_fn, 0.16, 1).update_fn(torch.tensor(xb), vt, None, None)
        raise SystemExit("expected the reference's Langevin corrector to fail on the sub-VP SDE (no `alphas`)")
    except AttributeError:
        pass
    out["noise_vp"] = dict(x=xb, t=np.float32(t_i), z0=zs[0], z1=zs[1], **nz)

    # ---- 6. IPO: RotOpt + Adam ----------------------------------------------------------------------
    print("IPO (RotOpt + Adam)")
    ipo = {}
    for tag, zcfg in (("h36m", zo.H36M_ZEDO_CFG), ("mini", zo.MINI_ZEDO_CFG)):
        B = 16
        uv = torch.tensor(uvc[:, :, :2])
        Kt = torch.tensor(Kc)
        x0 = zo.init_hypothesis(ds["clusters"], 1, B)
        pelvis = torch.cat((uv[:, 0, :], torch.ones((B, 1))), axis=-1)
        T0 = torch.inverse(Kt).bmm(pelvis[:, :, None]).permute(0, 2, 1)
        T0 = T0 / torch.norm(T0, dim=-1, keepdim=True) * zcfg["IPO_T"]
        ck.check(f"IPO[{tag}] T0", zo.init_translation(uvc[:, :, :2], Kc, zcfg["IPO_T"]), T0.numpy(), 1e-6)
        rot = R.szo.RotOpt(B, axis=zcfg["RotAxes"], minT=zcfg["IPO_minScaleT"], maxT=zcfg["IPO_maxScaleT"])
        opt = torch.optim.Adam(rot.parameters(), lr=0.1)
        crit = torch.nn.L1Loss(reduction="none")
        kl = zcfg["IPO_keylist"]
        xk = torch.tensor(x0)[:, kl, :]
        qs, ss, losses, grads = [], [], [], []
        mask = zo.axes_to_mask(zcfg["RotAxes"])

        def cur_q():
            cols = [rot.rot_vect.detach().numpy()]
            for a in "xyz":
                p = getattr(rot, f"rot_vect_{a}", None)
                cols.append(p.detach().numpy() if p is not None else np.zeros((B, 1), np.float32))
            return np.concatenate(cols, axis=1).astype(np.float32)

        n_it = 500
        for it in range(n_it):
            opt.zero_grad()
            rot2d = rot(xk, T0, Kt)
            loss = torch.mean(crit(rot2d[:, :, :2], uv[:, kl, :2]))
            loss.backward()
            if it < 60:
                q_now, s_now = cur_q(), rot.scale.detach().numpy().reshape(B).copy()
                l_o, dq_o, ds_o = zo.ipo_loss_and_grad(q_now, s_now, x0[:, kl, :], uvc[:, kl, :2], T0.numpy(),
                                                       Kc, zcfg["IPO_minScaleT"], zcfg["IPO_maxScaleT"], mask)
                gcols = [rot.rot_vect.grad.numpy()]
                for a in "xyz":
                    p = getattr(rot, f"rot_vect_{a}", None)
                    gcols.append(p.grad.numpy() if p is not None else np.zeros((B, 1), np.float32))
                g_ref = np.concatenate(gcols, axis=1)
                grads.append((rel_err(dq_o, g_ref), rel_err(ds_o, rot.scale.grad.numpy().reshape(B)),
                              abs(l_o - float(loss)) / float(loss)))
            opt.step()
            qs.append(cur_q())
            ss.append(rot.scale.detach().numpy().reshape(B).copy())
            losses.append(float(loss))
        ge = np.array(grads)
        ck.bound(f"IPO[{tag}] teacher-forced dL/dq rel err (60 iters, worst)", ge[:, 0].max(), 2e-5)
        ck.bound(f"IPO[{tag}] teacher-forced dL/dscale rel err (worst)", ge[:, 1].max(), 2e-5)
        ck.bound(f"IPO[{tag}] teacher-forced loss rel err (worst)", ge[:, 2].max(), 2e-6)
        trace = []
        Ro, To = zo.ipo_fit(x0, uvc[:, :, :2], Kc, kl, zcfg["RotAxes"], zcfg["IPO_T"], zcfg["IPO_minScaleT"],
                            zcfg["IPO_maxScaleT"], iters=n_it, trace=trace)
        ck.check(f"IPO[{tag}] free-running q after 10 iters", trace[9][0], qs[9], 2e-5)
        ck.check(f"IPO[{tag}] free-running scale after 10 iters", trace[9][1], ss[9], 2e-5)
        # long free-running trajectories are chaotic (L1 + Adam lr 0.1): report, compare the loss level
        drift = rel_err(trace[-1][0], qs[-1])
        uv_o, _, _ = zo.ipo_project(trace[-1][0], trace[-1][1], x0[:, kl, :], T0.numpy(), Kc,
                                    zcfg["IPO_minScaleT"], zcfg["IPO_maxScaleT"])
        loss_o = float(np.abs(uv_o - uvc[:, kl, :2]).mean())
        print(f"      (info) q drift after {n_it} free-running iterations: {drift:.2e}; "
              f"final L1 loss oracle {loss_o:.4f} vs reference {losses[-1]:.4f}")
        R_ref = rot.generate_matrix().detach().numpy()
        T_ref = (T0 * torch.clamp(rot.scale, min=zcfg["IPO_minScaleT"], max=zcfg["IPO_maxScaleT"])).detach().numpy()
        ck.check(f"IPO[{tag}] quaternion_to_matrix", zo.quaternion_to_matrix(qs[-1]), R_ref, 1e-6)
        ipo[tag] = dict(x0=x0, q_traj=np.stack(qs[:60]), s_traj=np.stack(ss[:60]), loss=np.array(losses),
                        R_final=R_ref, T_final=T_ref, T0=T0.numpy(), q_final=qs[-1], s_final=ss[-1])
    out["ipo"] = {f"{t}_{k}": v for t, d in ipo.items() for k, v in d.items()}

    # ---- 7. the full OIL loop of the driver, fed with the reference's (R, T) -------------------------
    print("OIL loop (1000 steps, B=16) -- takes a minute")
    B = 16
    R_ref, T_ref = ipo["h36m"]["R_final"], ipo["h36m"]["T_final"]
    x0 = ipo["h36m"]["x0"]
    uv_t = torch.tensor(uvc[:, :, :2])
    conf_t = torch.tensor(uvc[:, :, 2].copy())
    Kt = torch.tensor(Kc)
    sampling_fn = R.sampling.get_sampling_fn(cfg, sde, (B, 17, 3), lambda x: x, 0.01, device="cpu")
    with torch.no_grad():
        dx = torch.tensor(R_ref).bmm(torch.tensor(x0).permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        Tt = torch.tensor(T_ref)
        dumps_ref = {}
        n_steps = 1000
        for i in range(n_steps):
            if i < n_steps // 5:
                jg = R.szo.gradient_field_gen(uv_t, dx, Kt, t=Tt, conf=conf_t, returnT=False)
            else:
                jg, Tt = R.szo.gradient_field_gen(uv_t, dx, Kt, conf=conf_t, returnT=True)
            dx += jg
            _, results = sampling_fn(model, condition=uv_t * 0, gradient=jg, denoise_x=dx, t=ts[i], t_step=i,
                                     args=None)
            dx = torch.tensor(results)
            if i in (0, 9, 99, 199, 200, 299, 499, 799, 999):
                dumps_ref[i] = results.copy()
    x_rot = np.einsum("bij,bnj->bni", R_ref, x0).astype(np.float32)
    xo, To, dumps_o = zo.oil_loop(W, x_rot, T_ref, uvc[:, :, :2], Kc, uvc[:, :, 2].copy(), dump_every=1)
    dmap = dict(dumps_o)
    # The loop is expansive with random-init weights: two float32 implementations that agree to
    # ~1e-6 per step (numpy vs torch sgemm/inverse rounding) drift apart cumulatively.  That drift
    # is the float32 noise floor of the reference itself and is what "cumulative" parity of any
    # non-bit-identical implementation has to be read against; per-step parity is teacher-forced.
    for i in sorted(dumps_ref):
        ck.check(f"OIL cumulative pose after step {i} (fp32 noise floor)", dmap[i], dumps_ref[i], 3e-3)
    ck.check("OIL final T", To, Tt.numpy(), 3e-3)
    # teacher-forced: restart the oracle from the reference's state at the previous dump
    with torch.no_grad():
        xs = torch.tensor(dumps_ref[499]).clone()
        jg, Tn = R.szo.gradient_field_gen(uv_t, xs, Kt, conf=conf_t, returnT=True)
        xs += jg
        _, res500 = sampling_fn(model, condition=uv_t * 0, gradient=jg, denoise_x=xs, t=ts[500], t_step=500, args=None)
    g500, T500 = zo.gradient_field(uvc[:, :, :2], dumps_ref[499], Kc, conf=uvc[:, :, 2].copy())
    _, tf500 = zo.pc_sampler_step(W, (dumps_ref[499] + g500).astype(np.float32), np.float32(ts[500]))
    ck.check("OIL teacher-forced step 500 (from reference state)", tf500, res500, 1e-5)
    gt16 = ds["db_3d"].astype(np.float64)
    d_mpjpe = abs(np.mean([zo.mpjpe(xo[n], gt16[n]) for n in range(B)]) -
                  np.mean([zo.mpjpe(dumps_ref[999][n], gt16[n]) for n in range(B)]))
    ck.bound("OIL final MPJPE difference in metres (tol 0.1 mm)", d_mpjpe, 1e-4)
    out["tf500"] = dict(x_in=dumps_ref[499], t=np.float32(ts[500]), x_out=res500, T_out=Tn.numpy())
    out["oil"] = dict(R=R_ref, T=T_ref, x0=x0, T_final=Tt.numpy(),
                      steps=np.array(sorted(dumps_ref)), poses=np.stack([dumps_ref[i] for i in sorted(dumps_ref)]))

    # ---- 7b. the same loop with a damped network: poses stay at realistic scale (MPJPE ~ 0.1 m) ------------
    # Is the cumulative drift a property of the random-init network?  No: with post_dense scaled by 0.05 (network
    # push 20x smaller) two float32 implementations still end ~1e-3 apart.  The source is the per-step least-squares
    # translation (cond ~ 1e3 along the depth axis): the reference's float32 solve is 4.5e-6 off the exact solution
    # (2e-5 m), a random walk over 800 phase-2 steps; x is never re-centred, so depth noise lands in MPJPE directly.
    # Per-pose MPJPE of the reference is therefore only reproducible to ~1 mm across BLAS/LAPACK stacks; the
    # aggregate (mean over poses) is much tighter.
    print("OIL loop, damped network (1000 steps, B=16)")
    Ws = dict(W)
    Ws["post_dense.weight"] = (W["post_dense.weight"] * np.float32(0.05)).astype(np.float32)
    Ws["post_dense.bias"] = (W["post_dense.bias"] * np.float32(0.05)).astype(np.float32)
    model_s = load_into(torch, R.ScoreModelFC_Adv(cfg, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512,
                                                  cond_dim=3), Ws)
    conf_t = torch.tensor(uvc[:, :, 2].copy())
    with torch.no_grad():
        dx = torch.tensor(R_ref).bmm(torch.tensor(x0).permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        Tt = torch.tensor(T_ref)
        for i in range(n_steps):
            if i < n_steps // 5:
                jg = R.szo.gradient_field_gen(uv_t, dx, Kt, t=Tt, conf=conf_t, returnT=False)
            else:
                jg, Tt = R.szo.gradient_field_gen(uv_t, dx, Kt, conf=conf_t, returnT=True)
            dx += jg
            _, results = sampling_fn(model_s, condition=uv_t * 0, gradient=jg, denoise_x=dx, t=ts[i], t_step=i,
                                     args=None)
            dx = torch.tensor(results)
    xs, Ts, _ = zo.oil_loop(Ws, x_rot, T_ref, uvc[:, :, :2], Kc, uvc[:, :, 2].copy())
    ck.check("damped OIL final pose (cumulative)", xs, results, 3e-3)
    mp_ref = np.array([zo.mpjpe(results[n], gt16[n]) for n in range(B)])
    mp_or = np.array([zo.mpjpe(xs[n], gt16[n]) for n in range(B)])
    print(f"      (info) damped loop MPJPE level {mp_ref.mean():.4f} m")
    print(f"      (info) per-pose |dMPJPE| mean {np.abs(mp_ref - mp_or).mean() * 1e3:.3f} mm, max "
          f"{np.abs(mp_ref - mp_or).max() * 1e3:.3f} mm; aggregate {abs(mp_ref.mean() - mp_or.mean()) * 1e3:.4f} mm")
    ck.bound("damped OIL per-pose |dMPJPE| max in metres (float32 noise floor, tol 3 mm)",
             np.abs(mp_ref - mp_or).max(), 3e-3)
    ck.bound("damped OIL aggregate |dMPJPE| in metres (tol 0.1 mm per metre of MPJPE)",
             abs(mp_ref.mean() - mp_or.mean()), 1e-4 * max(1.0, mp_ref.mean()))
    out["oil_small"] = dict(x_final=results, T_final=Tt.numpy(), post_scale=np.float32(0.05), mpjpe=mp_ref)

    # ---- 7c. BASELINE configs[0] size: 1,024 poses through the reference's IPO (500 Adam iterations) and its
    # 1000-step OIL loop (damped network, realistic pose scale).  The aggregate MPJPE over 1,024 poses is the
    # quantity north_star bounds by 0.1 mm; per-pose values carry the float32 noise floor discussed above.
    print("C1 size: 1024 poses, reference IPO + OIL loop (1000 steps) -- takes a few minutes")
    N1, seed1 = 1024, 2024
    ds1 = zo.make_synthetic_dataset(N1, seed=seed1, n_clusters=1)
    uv1, K1n = ds1["db_2d"].copy(), ds1["camera_param"].copy()
    zcfg = zo.H36M_ZEDO_CFG
    x01 = zo.init_hypothesis(ds1["clusters"], 0, N1)
    uv1_t, K1_t = torch.tensor(uv1[:, :, :2]), torch.tensor(K1n)
    pelvis = torch.cat((uv1_t[:, 0, :], torch.ones((N1, 1))), axis=-1)
    T01 = torch.inverse(K1_t).bmm(pelvis[:, :, None]).permute(0, 2, 1)
    T01 = T01 / torch.norm(T01, dim=-1, keepdim=True) * zcfg["IPO_T"]
    rot1 = R.szo.RotOpt(N1, axis=zcfg["RotAxes"], minT=zcfg["IPO_minScaleT"], maxT=zcfg["IPO_maxScaleT"])
    opt1 = torch.optim.Adam(rot1.parameters(), lr=0.1)
    crit1 = torch.nn.L1Loss(reduction="none")
    kl1 = zcfg["IPO_keylist"]
    xk1 = torch.tensor(x01)[:, kl1, :]
    for it in range(zcfg["IPO_iterations"]):
        opt1.zero_grad()
        rot2d = rot1(xk1, T01, K1_t)
        loss = torch.mean(crit1(rot2d[:, :, :2], uv1_t[:, kl1, :2]))
        loss.backward()
        opt1.step()
    R1 = rot1.generate_matrix().detach().numpy()
    T1 = (T01 * torch.clamp(rot1.scale, min=zcfg["IPO_minScaleT"], max=zcfg["IPO_maxScaleT"])).detach().numpy()
    sampling_fn1 = R.sampling.get_sampling_fn(cfg, sde, (N1, 17, 3), lambda x: x, 0.01, device="cpu")
    conf1_t = torch.tensor(uv1[:, :, 2].copy())
    with torch.no_grad():
        dx = torch.tensor(R1).bmm(torch.tensor(x01).permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        Tt1 = torch.tensor(T1)
        for i in range(n_steps):
            if i < n_steps // 5:
                jg = R.szo.gradient_field_gen(uv1_t, dx, K1_t, t=Tt1, conf=conf1_t, returnT=False)
            else:
                jg, Tt1 = R.szo.gradient_field_gen(uv1_t, dx, K1_t, conf=conf1_t, returnT=True)
            dx += jg
            _, res1 = sampling_fn1(model_s, condition=uv1_t * 0, gradient=jg, denoise_x=dx, t=ts[i], t_step=i,
                                   args=None)
            dx = torch.tensor(res1)
    gt1 = ds1["db_3d"].astype(np.float64)
    mp1_ref = np.array([zo.mpjpe(res1[n], gt1[n]) for n in range(N1)])
    x_rot1 = np.einsum("bij,bnj->bni", R1, x01).astype(np.float32)
    xs1, _, _ = zo.oil_loop(Ws, x_rot1, T1, uv1[:, :, :2], K1n, uv1[:, :, 2].copy())
    mp1_or = np.array([zo.mpjpe(xs1[n], gt1[n]) for n in range(N1)])
    print(f"      (info) C1 MPJPE level {mp1_ref.mean():.4f} m; per-pose |dMPJPE| mean "
          f"{np.abs(mp1_ref - mp1_or).mean() * 1e3:.3f} mm, max {np.abs(mp1_ref - mp1_or).max() * 1e3:.3f} mm; "
          f"aggregate {abs(mp1_ref.mean() - mp1_or.mean()) * 1e3:.4f} mm")
    ck.bound("C1 (1024 poses) aggregate |dMPJPE| in metres (north_star: 0.1 mm)", abs(mp1_ref.mean() - mp1_or.mean()),
             1e-4)
    ck.bound("C1 (1024 poses) per-pose |dMPJPE| mean in metres (float32 noise floor, tol 1 mm)",
             np.abs(mp1_ref - mp1_or).mean(), 1e-3)
    out["c1"] = dict(seed=np.int64(seed1), R=R1, T=T1, x_final=res1, T_final=Tt1.numpy(), mpjpe=mp1_ref,
                     post_scale=np.float32(0.05))

    # ---- 8. Procrustes and eval_multi ---------------------------------------------------------------
    print("procrustes / eval_multi")
    rng = np.random.default_rng(11)
    N, S = 30, 5
    gts_mm = (ds["db_3d"][:1].astype(np.float64) * 0 + rng.normal(0, 300, (N, 17, 3)))
    gts_mm += rng.uniform(-500, 500, (N, 1, 3))
    gts = (gts_mm - gts_mm[:, 0:1]) / 1000.0
    preds = (gts[:, None] + rng.normal(0, 0.05, (N, S, 17, 3))).astype(np.float32)
    preds[3, 2] = (gts[3] * np.array([-1, 1, 1])).astype(np.float32)          # a reflected hypothesis
    preds[4, 1] = (1.7 * gts[4] @ np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])).astype(np.float32)
    preds[5, 0] = preds[5, 3]                                                  # exact tie -> first index wins
    al_ref = np.stack([R.transforms.align_to_gt(pose=preds[n, s], pose_gt=gts[n]) for n in range(N) for s in range(S)])
    al_o = np.stack([zo.procrustes_align(preds[n, s], gts[n]) for n in range(N) for s in range(S)])
    ck.check("align_to_gt (incl. reflected + scaled cases)", al_o, al_ref, 1e-9)
    items = [dict(joint_3d_camera=gts_mm[n], action=int(2 + n % 15)) for n in range(N)]
    h36m = R.H36M.__new__(R.H36M)
    h36m.subset, h36m.gt_dataset, h36m.seq5678 = "test", items, False
    actions = np.array([it["action"] for it in items])
    ev = {}
    for p2 in (False, True):
        with redirect_stdout(io.StringIO()):
            e_ref = h36m.eval_multi(preds, protocol2=p2, print_verbose=False)
        # argmin indices are internal to eval_multi: rebuild them with the reference's own align_to_gt
        idx_ref, res_ref = [], []
        for n in range(N):
            errs = []
            for s in range(S):
                p = preds[n, s]
                if p2:
                    p = R.transforms.align_to_gt(pose=p, pose_gt=gts[n])
                errs.append(np.mean(np.sqrt(np.square(p - gts[n]).sum(axis=1))))
            idx_ref.append(int(np.argmin(errs)))
            res_ref.append(np.amin(errs))
        agg, res, idx = zo.eval_multi(preds, gts, protocol2=p2, actions=actions)
        ck.check(f"eval_multi(protocol2={p2}) H36M aggregate", agg, e_ref, 1e-12)
        ck.check(f"eval_multi(protocol2={p2}) per-pose min", res, res_ref, 1e-12)
        assert list(idx) == idx_ref, "argmin indices differ"
        print(f"  [ok] eval_multi(protocol2={p2}) argmin indices bit-exact")
        ev[f"agg_p{int(p2)}"] = np.float64(e_ref)
        ev[f"min_p{int(p2)}"] = np.array(res_ref)
        ev[f"idx_p{int(p2)}"] = np.array(idx_ref)
    sel = np.array(ev["idx_p0"])
    min_pred = preds[np.arange(N), sel]
    pck_ref = R.mutils.compute_PCK(preds=min_pred.reshape((-1, 17, 3)), gts=gts)
    auc_ref = R.mutils.compute_AUC(preds=min_pred.reshape((-1, 17, 3)), gts=gts)
    ck.check("compute_PCK", zo.compute_pck(gts, min_pred), pck_ref, 1e-12)
    ck.check("compute_AUC", zo.compute_auc(gts, min_pred), auc_ref, 1e-12)
    ev["pck"], ev["auc"] = np.float64(pck_ref), np.float64(auc_ref)
    pw = R.PW3D.__new__(R.PW3D)
    pw.db_3d = gts
    with redirect_stdout(io.StringIO()):
        e_pw = pw.eval_multi(preds, protocol2=True)
    agg_pw, _, _ = zo.eval_multi(preds, gts, protocol2=True)
    ck.check("PW3D.eval_multi(protocol2=True) plain mean", agg_pw, e_pw, 1e-12)
    ev["agg_pw3d_p1"] = np.float64(e_pw)
    out["eval"] = dict(preds=preds, gts=gts, actions=actions, aligned=al_ref.astype(np.float64), **ev)

    pin_formats(R, ck, out)

    print()
    if not ck.all_ok():
        raise SystemExit("oracle does NOT match the reference")
    print(f"oracle matches the reference on {len(ck.rows)} checks")
    if not args.check:
        os.makedirs(GOLD, exist_ok=True)
        only = {n for n in args.only.split(",") if n}
        for name, d in out.items():
            if only and name not in only:
                continue
            path = os.path.join(GOLD, f"{name}.npz")
            np.savez_compressed(path, **d)
            print(f"wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.0f} KiB)")
        with open(os.path.join(GOLD, "PINNING.txt"), "w") as f:
            f.write("Golden vectors written by oracle/gen_golden.py from the imported reference "
                    f"(torch {torch.__version__}, CPU).\n")
            for name, e, tol, ok in ck.rows:
                f.write(f"{'ok  ' if ok else 'FAIL'} {name:60s} rel_err={e:.3e} tol={tol:.0e}\n")


if __name__ == "__main__":
    main()
