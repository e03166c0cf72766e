"""The reference-facing mirror (zedo_release_b200.lib.*) driven the way run/opt_main.py drives the
reference: same calls, same shapes, same return types -- checked against the golden vectors that
oracle/gen_golden.py recorded from the real reference."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import zedo_oracle as zo
from conftest import rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: there is no CPU fallback")
    import zedo_release_b200.lib as zlib
    zlib.install()  # `import lib...` now resolves to the mirror, as the reference drivers expect
    return zlib


def ref_config():
    return NS(training=NS(sde="subvpsde", continuous=True, cond_pose_mask_prob=0.0, cond_part_mask_prob=0.0,
                          cond_joint_mask_prob=0.0),
              sampling=NS(method="pc", predictor="euler_maruyama", corrector="none", snr=0.16, n_steps_each=1,
                          probability_flow=True, noise_removal=True),
              model=NS(embedding_type="positional", scale_by_sigma=False, sigma_max=50, sigma_min=0.01,
                       num_scales=1000, beta_min=0.1, beta_max=20.0, t=0.1, ema_rate=0.9999),
              device=torch.device("cuda"))


@pytest.fixture(scope="module")
def model(lib):
    from lib.algorithms.advanced.model import ScoreModelFC_Adv
    m = ScoreModelFC_Adv(ref_config(), n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    sd = {k: torch.tensor(v) for k, v in zo.make_weights(seed=0).items()}
    sd["sigmas"] = m.sigmas.clone()
    m.load_state_dict(sd)
    m.to(torch.device("cuda"))
    m.eval()
    return m


def test_model_forward_matches_reference(lib, model, golden):
    g = golden("net")
    x = torch.tensor(g["x"], device="cuda")
    for t in (0.1, 0.05, 0.01):
        out = model(x, torch.ones(8, device="cuda") * torch.tensor(t) * 999, torch.zeros(8, 17, 2, device="cuda"), None)
        assert out.shape == (8, 17, 3) and rel_err(out.cpu().numpy(), g[f"out_{t}"]) < 2e-5
    # per-row labels (never used by the drivers) still give the right rows
    lab = torch.tensor([99.9] * 4 + [9.99] * 4, device="cuda")
    out = model(x, lab, None, None).cpu().numpy()
    assert rel_err(out[:4], g["out_0.1"][:4]) < 2e-5 and rel_err(out[4:], g["out_0.01"][4:]) < 2e-5
    # a weight update invalidates the packed plan
    with torch.no_grad():
        model.post_dense.bias.add_(1.0)
    out2 = model(x, torch.ones(8, device="cuda") * 99.9, None, None).cpu().numpy()
    assert rel_err(out2, g["out_0.1"] + 1.0) < 2e-5
    with torch.no_grad():
        model.post_dense.bias.sub_(1.0)
    # the mirror's EMA writes through the parameters (version counters move): copy_to / restore re-pack the plan
    from lib.algorithms.ema import ExponentialMovingAverage
    ema = ExponentialMovingAverage(model.parameters(), decay=0.999)
    with torch.no_grad():
        ema.shadow_params[-1].add_(2.0)  # post_dense.bias is the last parameter
    ema.store(model.parameters())
    ema.copy_to(model.parameters())
    out3 = model(x, torch.ones(8, device="cuda") * 99.9, None, None).cpu().numpy()
    assert rel_err(out3, g["out_0.1"] + 2.0) < 2e-5
    ema.restore(model.parameters())
    out4 = model(x, torch.ones(8, device="cuda") * 99.9, None, None).cpu().numpy()
    assert rel_err(out4, g["out_0.1"]) < 2e-5
    # a write through .data is invisible to the cache key: invalidate_plan() is the documented hook
    model.post_dense.bias.data.add_(1.0)
    model.invalidate_plan()
    out5 = model(x, torch.ones(8, device="cuda") * 99.9, None, None).cpu().numpy()
    assert rel_err(out5, g["out_0.1"] + 1.0) < 2e-5
    model.post_dense.bias.data.sub_(1.0)
    model.invalidate_plan()


def test_driver_loop_through_the_mirror(lib, model, golden):
    """run/opt_main.py:197-222 verbatim (100 of the 1000 steps), fed with the reference's (R, T)."""
    from lib.algorithms.advanced import sde_lib, sampling
    from lib.algorithms.advanced.simple_zeroshot_opt import gradient_field_gen
    g, geo = golden("oil"), golden("geom")
    config, device = ref_config(), torch.device("cuda")
    sde = sde_lib.subVPSDE(beta_min=config.model.beta_min, beta_max=config.model.beta_max,
                           N=config.model.num_scales, T=config.model.t)
    config.sampling.probability_flow = True
    sampling_fn = sampling.get_sampling_fn(config, sde, (16, 17, 3), lambda x: x, 0.01, device=device)
    condition = torch.tensor(geo["db_2d"][:, :, :2], device=device).float()
    conf = torch.tensor(geo["db_2d"][:, :, 2], device=device).float()
    K = torch.tensor(geo["K"], device=device).float()
    T = torch.tensor(g["T"], device=device)
    rot_mat = torch.tensor(g["R"], device=device)
    denoise_x = torch.tensor(g["x0"], device=device)
    sample_num = 1000
    timestamp = torch.linspace(sde.T, 0.01, sample_num, device=device)
    steps = list(g["steps"])
    with torch.no_grad():
        denoise_x = rot_mat.bmm(denoise_x.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        for i in range(0, 100):
            if i < sample_num // 5:
                joint_gradient = gradient_field_gen(condition, denoise_x, K, t=T, conf=conf, returnT=False)
            else:
                joint_gradient, T = gradient_field_gen(condition, denoise_x, K, conf=conf, returnT=True)
            denoise_x += joint_gradient
            trajs, results = sampling_fn(model, condition=condition * 0, gradient=joint_gradient,
                                         denoise_x=denoise_x, t=timestamp[i], t_step=i, args=None)
            assert isinstance(results, np.ndarray) and results.dtype == np.float32 and results.shape == (16, 17, 3)
            assert trajs.shape == (1, 16, 17, 3) and np.array_equal(trajs[0], results)
            denoise_x = torch.tensor(results).to(device)
            if i in (0, 9, 99):
                tol = {0: 5e-6, 9: 5e-5, 99: 5e-4}[i]
                assert rel_err(results, g["poses"][steps.index(i)]) < tol, i


def test_noise_bearing_predictors_through_the_mirror(lib, model, golden, monkeypatch):
    from lib.algorithms.advanced import sde_lib, sampling
    g, n = golden("sampler"), golden("noise")
    z = torch.tensor(n["z"], device="cuda")
    monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: z)  # the "identical injected noise tensors"
    sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
    x, t = torch.tensor(g["x"], device="cuda"), torch.tensor(float(g["t"]), device="cuda")
    for pred, kx, km in (("euler_maruyama", "em_x", "em_mean"), ("reverse_diffusion", "rd_x", "rd_mean")):
        fn = sampling.get_pc_sampler(sde, (8, 17, 3), sampling.get_predictor(pred), sampling.get_corrector("none"),
                                     lambda v: v, 0.16, probability_flow=False, continuous=True, denoise=False)
        trajs, res = fn(model, condition=torch.zeros(8, 17, 2, device="cuda"), denoise_x=x, t=t, t_step=0)
        assert rel_err(res, n[kx]) < 1e-5 and rel_err(trajs[0], n[km]) < 1e-5
        # the generic (unfused) composition of the registered classes agrees with the fused kernel
        score_fn = sampling.mutils.get_score_fn(sde, model, train=False, continuous=True)
        xs, xm = sampling.get_predictor(pred)(sde, score_fn, False).update_fn(x, torch.ones(8, device="cuda") * t,
                                                                             None, None)
        assert rel_err(xs.cpu().numpy(), n[kx]) < 1e-5 and rel_err(xm.cpu().numpy(), n[km]) < 1e-5


def test_ancestral_langevin_ald_fused_with_injected_noise(lib, model, golden, monkeypatch):
    """The remaining registered updates (sampling.py:208-324) through the mirror's classes: network forward + ONE fused
    update kernel each (zedo_score_stats / zedo_noise_update), the reference's recorded outputs for the same injected
    noise tensors as golden (VPSDE / VESDE; the reference's Langevin raises AttributeError on the sub-VP SDE)."""
    from lib.algorithms.advanced import sde_lib, sampling
    import zedo_release_b200 as zr
    g = golden("noise_vp")
    zs = [torch.tensor(g["z0"], device="cuda"), torch.tensor(g["z1"], device="cuda")]
    it = iter(())

    def fake_randn_like(x, *a, **k):
        return next(it)

    monkeypatch.setattr(torch, "randn_like", fake_randn_like)
    vp = sde_lib.VPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
    ve = sde_lib.VESDE(sigma_min=0.01, sigma_max=50.0, N=1000, T=0.1)
    sf_vp = sampling.mutils.get_score_fn(vp, model, train=False, continuous=True)
    sf_ve = sampling.mutils.get_score_fn(ve, model, train=False, continuous=True)
    x, vt = torch.tensor(g["x"], device="cuda"), torch.ones(8, device="cuda") * float(g["t"])
    cases = (("anc_vp", lambda: sampling.AncestralSamplingPredictor(vp, sf_vp)),
             ("anc_ve", lambda: sampling.AncestralSamplingPredictor(ve, sf_ve)),
             ("lang", lambda: sampling.LangevinCorrector(vp, sf_vp, 0.16, 2)),
             ("ald", lambda: sampling.AnnealedLangevinDynamics(vp, sf_vp, 0.16, 1)))
    n0 = zr._native.launch_count()
    for tag, make in cases:
        it = iter(zs)
        xs, xm = make().update_fn(x, vt, None, None)
        assert rel_err(xs.cpu().numpy(), g[f"{tag}_x"]) < 2e-5 and rel_err(xm.cpu().numpy(), g[f"{tag}_mean"]) < 2e-5, tag
    assert zr._native.launch_count() > n0  # the fused path ran (no eager composition)
    # non-uniform time labels fall back to the eager composition of the same classes and agree with it
    it = iter(zs)
    vt2 = vt.clone()
    vt2[4:] *= 0.5
    xs2, _ = sampling.AncestralSamplingPredictor(vp, sf_vp).update_fn(x, vt2, None, None)
    assert rel_err(xs2[:4].cpu().numpy(), g["anc_vp_x"][:4]) < 2e-5
    with pytest.raises(AttributeError):  # like the reference: subVPSDE has no `alphas`
        sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
        sampling.LangevinCorrector(sde, sampling.mutils.get_score_fn(sde, model, continuous=True), 0.16, 1).update_fn(
            x, vt, None, None)


def test_rotopt_adam_loop_through_the_mirror(lib, golden):
    """run/opt_main.py:180-195 verbatim: RotOpt + torch.optim.Adam + L1Loss, 10 iterations."""
    from lib.algorithms.advanced.simple_zeroshot_opt import RotOpt
    g, geo = golden("ipo"), golden("geom")
    device = torch.device("cuda")
    for tag, cfg in (("h36m", zo.H36M_ZEDO_CFG), ("mini", zo.MINI_ZEDO_CFG)):
        condition = torch.tensor(geo["db_2d"][:, :, :2], device=device).float()
        K = torch.tensor(geo["K"], device=device).float()
        denoise_x = torch.tensor(g[f"{tag}_x0"], device=device)
        pelvis = torch.cat((condition[:, 0, :], torch.ones((16, 1), device=device)), axis=-1)
        T = torch.inverse(K).bmm(pelvis[:, :, None]).permute(0, 2, 1)
        T = T / torch.norm(T, dim=-1, keepdim=True) * cfg["IPO_T"]
        rot_opt = RotOpt(16, axis=cfg["RotAxes"], minT=cfg["IPO_minScaleT"], maxT=cfg["IPO_maxScaleT"])
        rot_opt.to(device)
        opt = torch.optim.Adam(rot_opt.parameters(), lr=0.1)
        criterion = torch.nn.L1Loss(reduction='none')
        kl = cfg["IPO_keylist"]
        for i in range(10):
            opt.zero_grad()
            rot2d = rot_opt(denoise_x[:, kl, :], T, K)
            loss = torch.mean(criterion(rot2d[:, :, :2], condition[:, kl, :2]))
            loss.backward()
            opt.step()
            assert abs(float(loss.detach()) - g[f"{tag}_loss"][i]) / g[f"{tag}_loss"][i] < 1e-4
        q = torch.cat([rot_opt.rot_vect] + [getattr(rot_opt, f"rot_vect_{a}", torch.zeros(16, 1, device=device))
                                            for a in "xyz"], dim=-1).detach().cpu().numpy()
        assert rel_err(q, g[f"{tag}_q_traj"][9]) < 1e-4
        assert rel_err(rot_opt.scale.detach().cpu().numpy().reshape(16), g[f"{tag}_s_traj"][9]) < 1e-4
        assert rot_opt.generate_matrix().shape == (16, 3, 3)


def test_dataset_eval_multi_and_align_to_gt(lib, golden):
    from lib.dataset.synthetic import ArrayPoseDataset
    from lib.utils.transforms import align_to_gt
    g = golden("eval")
    ds = ArrayPoseDataset(g["gts"], np.zeros((30, 17, 3), np.float32), np.zeros((30, 3, 3), np.float32),
                          actions=g["actions"])
    for p2 in (False, True):
        e = ds.eval_multi(g["preds"], protocol2=p2, print_verbose=False)
        assert abs(e - float(g[f"agg_p{int(p2)}"])) < 2e-7
        assert np.array_equal(ds.last_index, g[f"idx_p{int(p2)}"])
    plain = ArrayPoseDataset(g["gts"], np.zeros((30, 17, 3), np.float32), np.zeros((30, 3, 3), np.float32))
    assert abs(plain.eval_multi(g["preds"], protocol2=True) - float(g["agg_pw3d_p1"])) < 2e-7
    for k in (0, 17, 149):  # incl. the reflected hypothesis (pose 3, hypothesis 2)
        n, s = divmod(k, 5)
        assert np.abs(align_to_gt(g["preds"][n, s], g["gts"][n]) - g["aligned"][k]).max() < 5e-7


def test_procrustes_returns_the_reference_tform(lib, golden):
    """``procrustes(A, B)`` = (d, Z, tform) of transforms.py:42-128: Z from the device kernel, tform such that
    Z = scale * B @ rotation + translation; checked against the reference's formula (numpy SVD) on a plain and on the
    reflected hypothesis of the golden set."""
    from lib.utils.transforms import procrustes
    g = golden("eval")
    for k in (0, 17):
        n, s = divmod(k, 5)
        A, B = g["gts"][n].astype(np.float64), g["preds"][n, s].astype(np.float64)
        d, Z, tf = procrustes(A, B)
        A0, B0 = A - A.mean(0), B - B.mean(0)
        an, bn = np.sqrt((A0 ** 2).sum()), np.sqrt((B0 ** 2).sum())
        U, sv, Vt = np.linalg.svd((A0 / an).T @ (B0 / bn))
        R = Vt.T @ U.T
        scale = sv.sum() * an / bn
        assert np.abs(Z - g["aligned"][k]).max() < 5e-7
        assert abs(d - (1 - sv.sum() ** 2)) < 1e-6
        assert np.abs(tf["rotation"] - R).max() < 1e-5 and abs(tf["scale"] - scale) < 1e-5 * scale
        assert np.abs(tf["translation"] - (A.mean(0) - scale * B.mean(0) @ R)).max() < 1e-5
        assert np.abs(tf["scale"] * B @ tf["rotation"] + tf["translation"] - Z).max() < 1e-6
    try:  # not on the evaluation path: the reference's own function when a checkout is known to the mirror, else an error
        _, Z1, tf1 = procrustes(A, B, scaling=False)
    except NotImplementedError:
        pass
    else:
        assert tf1["scale"] == 1 and np.abs(np.sqrt((B0 ** 2).sum()) - np.sqrt(((Z1 - Z1.mean(0)) ** 2).sum())) < 1e-9


def test_compute_pck_auc_have_the_reference_signature(lib, golden):
    """``lib.algorithms.advanced.utils.compute_PCK / compute_AUC`` (utils.py:814-849, imported by the reference's
    MPII3DHP loader) on the device: equal to the reference's formula -- strict ``<`` on errors in millimetres, all joints
    or ``eval_joints``, AUC = mean over linspace(0, 150, 31)."""
    from lib.algorithms.advanced.utils import compute_AUC, compute_PCK
    g = golden("eval")
    gts = (g["gts"] - g["gts"][:, 0:1]).astype(np.float64)
    preds = g["preds"][:, 1].astype(np.float32)
    err_mm = np.sqrt(((preds.astype(np.float64) - gts) ** 2).sum(-1)) * 1000.0   # [N, J]
    for joints in (None, [1, 2, 3, 14, 15, 16]):
        e = err_mm if joints is None else err_mm[:, joints]
        want = [100.0 * float((e < t).sum()) / e.size for t in np.linspace(0, 150, 31)]
        assert abs(compute_PCK(gts, preds, eval_joints=joints) - want[30]) < 1e-9
        assert abs(compute_PCK(gts, preds, eval_joints=joints, threshold=50) - want[10]) < 1e-9
        assert abs(compute_AUC(gts, preds, eval_joints=joints) - float(np.mean(want))) < 1e-9
    with pytest.raises(NotImplementedError):
        compute_PCK(gts, preds, threshold=42)


def test_dataset_eval_variants(lib, golden):
    """valid_ind filtering, the literal sample_interval semantics and the 3DHP extras (PCK / AUC / std)."""
    from lib.dataset.synthetic import ArrayPoseDataset
    g = golden("eval")
    preds, gts = g["preds"], g["gts"]
    N, S = preds.shape[:2]
    zeros2d, zerosK = np.zeros((N, 17, 3), np.float32), np.zeros((N, 3, 3), np.float32)
    rng = np.random.default_rng(5)
    valid = [sorted(rng.choice(S, size=rng.integers(1, S + 1), replace=False).tolist()) for _ in range(N)]
    plain = ArrayPoseDataset(gts, zeros2d, zerosK)
    h36m = ArrayPoseDataset(gts, zeros2d, zerosK, actions=g["actions"])
    for p2 in (False, True):
        agg, res, idx = zo.eval_multi(preds, gts, protocol2=p2, valid_ind=valid)
        assert abs(plain.eval_multi(preds, protocol2=p2, valid_ind=valid) - agg) < 2e-7
        assert np.array_equal(plain.last_index, idx)
        agg_a, _, _ = zo.eval_multi(preds, gts, protocol2=p2, actions=g["actions"], valid_ind=valid)
        assert abs(h36m.eval_multi(preds, protocol2=p2, valid_ind=valid) - agg_a) < 2e-7
    with pytest.raises(ValueError):
        plain.eval_multi(preds, valid_ind=[[]] * N)
    # sample_interval: preds[::k] against the first len(preds[::k]) ground truths, as the reference does
    agg3, _, idx3 = zo.eval_multi(preds[::3], gts[:len(preds[::3])])
    assert abs(plain.eval_multi(preds, sample_interval=3) - agg3) < 2e-7 and np.array_equal(plain.last_index, idx3)
    with pytest.raises(IndexError):
        h36m.eval_multi(preds, sample_interval=3)
    assert abs(h36m.eval_multi(preds, sample_interval=1) - float(g["agg_p0"])) < 2e-7
    # 3DHP: PCK / AUC of the selected (unaligned) hypotheses and the diversity std
    hp = ArrayPoseDataset(gts, zeros2d, zerosK, name="3dhp")
    hp.eval_multi(preds, protocol2=False)
    _, _, idx = zo.eval_multi(preds, gts)
    best = preds[np.arange(N), idx]
    assert abs(hp.last_pck - zo.compute_pck(gts, best)) < 1e-9 and abs(hp.last_auc - zo.compute_auc(gts, best)) < 1e-9
    assert np.allclose(hp.last_std, zo.hypothesis_std(preds.astype(np.float64)), rtol=1e-9)
    assert np.allclose(hp.last_std, zo.hypothesis_std(preds), rtol=1e-4)  # the reference's float32 arithmetic


def test_dataset_formats_against_the_reference_loaders(lib, golden):
    """ArrayPoseDataset.from_h36m_items / from_pw3d_npz build the arrays the reference's own loaders built from the
    same synthetic files (tests/golden/formats.npz), and eval_multi on them gives the reference's numbers."""
    from lib.dataset.synthetic import ArrayPoseDataset
    g = golden("formats")
    ds = zo.make_synthetic_dataset(30, seed=int(g["seed"]), n_clusters=2)
    items = zo.h36m_items_from_arrays(ds)
    h = ArrayPoseDataset.from_h36m_items(items, gt2d=True, abs_coord=True)
    assert np.array_equal(h.db_3d, g["h36m_db3d"]) and np.array_equal(h.camera_param, g["h36m_K"])
    assert np.array_equal(h.db_2d, g["h36m_db2d_gt"].astype(np.float32))
    hd = ArrayPoseDataset.from_h36m_items(items, gt2d=False, detections=(g["det_xy"], g["det_conf"]))
    assert np.array_equal(hd.db_2d, g["h36m_db2d_dt"])
    with pytest.raises(ValueError):
        ArrayPoseDataset.from_h36m_items(items, gt2d=False)
    assert abs(h.eval_multi(g["preds"], protocol2=False) - float(g["h36m_eval_p1"])) < 2e-7
    assert abs(h.eval_multi(g["preds"], protocol2=True) - float(g["h36m_eval_p2"])) < 2e-7
    pw = ArrayPoseDataset.from_pw3d_npz(zo.pw3d_npz_from_arrays(ds), abs_coord=True)
    assert np.array_equal(pw.db_3d, g["pw3d_db3d"]) and np.array_equal(pw.camera_param, g["pw3d_K"])
    assert np.abs(pw.db_2d - g["pw3d_db2d"]).max() < 1e-3  # pixels; einsum vs per-pose dot
    assert abs(pw.eval_multi(g["preds"], protocol2=False) - float(g["pw3d_eval_p1"])) < 2e-7


def test_control_model_through_the_mirror(lib, golden):
    from lib.algorithms.advanced.control_model import Control_ScoreModelFC_Adv
    from lib.algorithms.advanced import utils as mutils, sde_lib
    g = golden("control")
    W = zo.make_weights(seed=int(g["weights_seed"]), control=True)
    m = Control_ScoreModelFC_Adv(ref_config(), n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    assert set(m.state_dict().keys()) == set(W.keys()) | {"sigmas"}
    sd = {k: torch.tensor(v) for k, v in W.items()}
    sd["sigmas"] = m.sigmas.clone()
    m.load_state_dict(sd)
    m.to(torch.device("cuda")).eval()
    x = torch.tensor(g["x"], device="cuda")
    out = m(x, torch.ones(8, device="cuda") * float(g["t999"]))
    assert rel_err(out.cpu().numpy(), g["out"]) < 2e-5
    # callable through get_model_fn (4 arguments), which the shipped reference class is not
    sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20.0, N=1000, T=0.1)
    score = mutils.get_score_fn(sde, m, train=False, continuous=True)(x, torch.ones(8, device="cuda") * 0.05, None, None)
    ref = -zo.control_score_forward(W, g["x"], np.float32(0.05) * np.float32(999)) / zo.subvp_marginal_std(np.float32(0.05))
    assert rel_err(score.cpu().numpy(), ref) < 1e-4


def test_fourier_embedding_through_the_mirror(lib, golden):
    """embedding_type = 'fourier' (configs/default_pose_gen_configs.py:71): same state_dict keys as the reference
    (gauss_proj.W), forward against the reference's outputs, scale_by_sigma divides by t (model.py:248,294-296)."""
    from lib.algorithms.advanced.model import ScoreModelFC_Adv
    g = golden("net_fourier")
    cfg = ref_config()
    cfg.model.embedding_type = "fourier"
    W = zo.make_weights(seed=int(g["weights_seed"]), fourier=True)
    m = ScoreModelFC_Adv(cfg, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    assert set(m.state_dict().keys()) == set(W.keys()) | {"sigmas"}
    sd = {k: torch.tensor(v) for k, v in W.items()}
    sd["sigmas"] = m.sigmas.clone()
    m.load_state_dict(sd)
    m.to(torch.device("cuda")).eval()
    x = torch.tensor(g["x"], device="cuda")
    for t in (0.1, 0.05, 0.01):
        lab = torch.ones(8, device="cuda") * torch.tensor(t) * 999
        out = m(x, lab, torch.zeros(8, 17, 2, device="cuda"), None)
        assert rel_err(out.cpu().numpy(), g[f"out_{t}"]) < 2e-5
        emb = m.gauss_proj(torch.log(lab[:1])).cpu().numpy()
        assert rel_err(emb, g[f"emb_{t}"]) < 3e-4  # torch's CUDA logf / sinf, 1 ulp of log t = 2e-4 here
    cfg.model.scale_by_sigma = True
    lab = torch.ones(8, device="cuda") * 49.95
    out = m(x, lab, None, None)
    assert rel_err(out.cpu().numpy(), g["out_0.05"] / np.float32(49.95)) < 2e-5
    cfg.model.scale_by_sigma = False


def test_checkpoint_file_in_the_reference_format(lib, golden, tmp_path):
    """run/opt_main.py:120-137 verbatim: torch.load -> strip `module.` -> model.load_state_dict ->
    ema.load_state_dict -> state['step']; the file is written the way the reference's trainer saves it
    (DataParallel-prefixed keys, EMA shadow parameters, step)."""
    import copy
    import os
    from lib.algorithms.advanced.model import ScoreModelFC_Adv
    from lib.algorithms.ema import ExponentialMovingAverage
    config = ref_config()
    trained = {k: torch.tensor(v) for k, v in zo.make_weights(seed=0).items()}
    src = ScoreModelFC_Adv(config, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    trained["sigmas"] = src.sigmas.clone()
    src.load_state_dict(trained)
    src_ema = ExponentialMovingAverage(src.parameters(), decay=config.model.ema_rate)
    src_ema.update(src.parameters())
    ckpt_path = os.path.join(tmp_path, "checkpoint_1500.pth")
    torch.save({"model_state_dict": {"module." + k: v for k, v in src.state_dict().items()},
                "ema": src_ema.state_dict(), "step": 1500, "optimizer": {}}, ckpt_path)

    model = ScoreModelFC_Adv(config, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    model.to(config.device)
    ema = ExponentialMovingAverage(model.parameters(), decay=config.model.ema_rate)
    state = dict(optimizer=None, model=model, ema=ema, step=0)
    old_checkpoint = torch.load(ckpt_path, map_location={'cuda:0': 'cuda:0'})
    checkpoint = copy.deepcopy(old_checkpoint)
    checkpoint['model_state_dict'] = {}
    for k, v in old_checkpoint['model_state_dict'].items():
        name = k[7:]  # remove `module.`
        checkpoint['model_state_dict'][name] = v
    model.load_state_dict(checkpoint['model_state_dict'])
    ema.load_state_dict(checkpoint['ema'])
    state['step'] = checkpoint['step']
    model.eval()

    assert state['step'] == 1500 and ema.num_updates == 1 and ema.decay == config.model.ema_rate
    assert len(ema.shadow_params) == len(list(model.parameters()))
    g = golden("net")
    out = model(torch.tensor(g["x"], device="cuda"), torch.ones(8, device="cuda") * 99.9, None, None)
    assert rel_err(out.cpu().numpy(), g["out_0.1"]) < 2e-5
    # the packed plan also accepts the prefixed names directly (C ABI: zedo_plan_create strips `module.`)
    import zedo_release_b200 as zr
    plan = zr.ScorePlan({k: v.cuda() for k, v in old_checkpoint['model_state_dict'].items()}, n_joints=17, max_batch=8)
    out2 = plan.forward(torch.tensor(g["x"], device="cuda"), 99.9)
    plan.close()
    assert torch.equal(out2, out)


def test_custom_dataset_read_data_and_eval(lib, tmp_path):
    """lib/dataset/custom.py: the reference's TODO ``read_data`` filled in (npz with 2D + confidence, optional 3D, K,
    image names), same constructor / attributes / eval_multi contract as run/inference.py:118-127,236-237 uses; the
    printed numbers equal the reference's own CustomDataset.eval_multi on the same arrays (numpy restatement)."""
    from lib.dataset.custom import CustomDataset
    ds = zo.make_synthetic_dataset(40, seed=21, n_clusters=2)
    np.savez(tmp_path / "custom.npz", keypoints_2d=ds["db_2d"], keypoints_3d=ds["db_3d"] + ds["root"][:, None],
             camera_params=ds["camera_param"])
    d = CustomDataset(str(tmp_path), sample_interval=2)
    assert len(d) == 20 and d.db_2d.shape == (20, 17, 3) and d.camera_param.shape == (20, 3, 3) and len(d.image_name) == 20
    rng = np.random.default_rng(0)
    preds = ((d.db_3d - d.db_3d[:, 0:1])[:, None] + rng.normal(0, 0.02, (20, 3, 17, 3))).astype(np.float32)
    for p2 in (False, True):
        got = d.eval_multi(preds, protocol2=p2, print_verbose=True)
        _, e_or, _ = zo.eval_multi(preds.astype(np.float64), (d.db_3d - d.db_3d[:, 0:1]).astype(np.float64), protocol2=p2)
        assert abs(got - float(np.mean(e_or))) < 1e-7
    # inference-only use: no 3D labels, a single K, no confidence column
    w = CustomDataset.from_arrays(ds["db_2d"][:, :, :2], ds["camera_param"][0])
    assert w.db_3d.shape == (40, 17, 3) and not w.db_3d.any() and np.all(w.db_2d[:, :, 2] == 1)
    with pytest.raises(ValueError):
        CustomDataset.from_arrays(ds["db_2d"][:, :, :1], ds["camera_param"])
