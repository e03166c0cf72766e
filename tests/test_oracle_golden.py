"""The oracle (oracle/zedo_oracle.py) against the golden vectors that oracle/gen_golden.py wrote
from the imported reference.  Runs without the reference tree and without a GPU."""
import numpy as np
import pytest

import zedo_oracle as zo
from conftest import rel_err


def test_time_grid_and_sde_scalars(golden):
    g = golden("sde")
    assert np.array_equal(zo.oil_time_grid(), g["time_grid"])  # bit-exact torch.linspace
    beta, diff = zo.subvp_sde_scalars(g["t"])
    assert rel_err(diff, g["diffusion"]) < 2e-5
    assert rel_err(0.5 * beta, g["half_beta"]) < 2e-7
    assert rel_err(zo.subvp_marginal_std(g["t"]), g["std"]) < 1e-4  # 1 - exp(.) cancellation, see _exp32
    # SURVEY 8c probe values
    for t, g2, sig in ((0.1, 0.411058, 0.103718), (0.05, 0.063510, 0.029433), (0.01, 0.00119064, 0.00199300)):
        _, d = zo.subvp_sde_scalars(np.float32(t))
        assert abs(float(d) ** 2 - g2) / g2 < 1e-4
        assert abs(float(zo.subvp_marginal_std(np.float32(t))) - sig) / sig < 1e-4


def test_reference_demo_known_answer(golden):
    """The only known-answer test the reference ships: simple_zeroshot_opt.py:127-147."""
    g = golden("demo")
    k3 = g["key3d"].copy()
    first = None
    for i in range(10):
        grad, _ = zo.gradient_field(g["key2d"], k3, g["K"])
        if i == 0:
            first = float(np.mean(np.linalg.norm(grad, axis=-1)))
        k3 = k3 + grad
    assert float(g["first_norm"]) == 53.63671875
    assert abs(first - 53.63671875) / 53.63671875 < 1e-6
    assert rel_err(k3, g["final_key3d"]) < 1e-6


@pytest.mark.parametrize("tag,use_t,use_conf", [("fixedT_conf", True, True), ("solveT_conf", False, True),
                                               ("solveT_noconf", False, False), ("fixedT_noconf", True, False)])
def test_gradient_field(golden, tag, use_t, use_conf):
    g = golden("geom")
    conf = g["db_2d"][:, :, 2].copy() if use_conf else None
    grad, T = zo.gradient_field(g["db_2d"][:, :, :2], g["x"], g["K"], t=g["T_in"] if use_t else None, conf=conf)
    assert rel_err(grad, g[f"{tag}_g"]) < 2e-5
    assert rel_err(T, g[f"{tag}_T"]) < 2e-5
    if use_conf:
        assert conf.max() <= 1.0 and conf.min() >= np.float32(1e-4)  # clamped in place


def test_gradient_field_sign_flip(golden):
    g = golden("geom")
    grad, T = zo.gradient_field(g["db_2d"][:, :, :2], g["x_neg"], g["K"])
    assert rel_err(grad, g["flip_g"]) < 1e-4 and rel_err(T, g["flip_T"]) < 2e-5
    assert (T[:, :, 2] >= 0).all()


def test_score_network(golden):
    g = golden("net")
    W = zo.make_weights(seed=int(g["weights_seed"]))
    for t in (0.1, 0.05, 0.01):
        out = zo.score_forward(W, g["x"], np.float32(t) * np.float32(999))
        assert rel_err(out, g[f"out_{t}"]) < 2e-5
    g12 = golden("net12")
    W12 = zo.make_weights(seed=int(g12["weights_seed"]), n_joints=12)
    assert rel_err(zo.score_forward(W12, g12["x"], g12["t999"]), g12["out"]) < 2e-5


def test_score_network_fourier_embedding(golden):
    """'fourier' time embedding (model.py:27-36,246-250; the default of configs/default_pose_gen_configs.py:71)."""
    g = golden("net_fourier")
    W = zo.make_weights(seed=int(g["weights_seed"]), fourier=True)
    Wp = zo.make_weights(seed=int(g["weights_seed"]))
    assert all(np.array_equal(W[k], Wp[k]) for k in Wp) and W["gauss_proj.W"].shape == (256,)
    for t in (0.1, 0.05, 0.01):
        t999 = np.float32(t) * np.float32(999)
        assert rel_err(zo.gaussian_fourier_projection(zo.log_f32(t999), W["gauss_proj.W"]), g[f"emb_{t}"]) < 2e-6
        assert rel_err(zo.score_forward(W, g["x"], t999), g[f"out_{t}"]) < 2e-5


def test_control_network(golden):
    g = golden("control")
    W = zo.make_weights(seed=int(g["weights_seed"]), control=True)
    assert rel_err(zo.control_score_forward(W, g["x"], g["t999"]), g["out"]) < 2e-5


def test_sampler_steps(golden):
    g, n = golden("sampler"), golden("noise")
    W = zo.make_weights(seed=0)
    trajs, res = zo.pc_sampler_step(W, g["x"], g["t"])
    assert rel_err(res, g["results"]) < 2e-6 and rel_err(trajs, g["trajs"]) < 2e-6
    x, xm = zo.euler_maruyama_update(W, g["x"], g["t"], z=n["z"], probability_flow=False)
    assert rel_err(x, n["em_x"]) < 2e-6 and rel_err(xm, n["em_mean"]) < 2e-6
    x, xm = zo.reverse_diffusion_update(W, g["x"], g["t"], z=n["z"], probability_flow=False)
    assert rel_err(x, n["rd_x"]) < 2e-6 and rel_err(xm, n["rd_mean"]) < 2e-6


@pytest.mark.parametrize("tag,cfg", [("h36m", zo.H36M_ZEDO_CFG), ("mini", zo.MINI_ZEDO_CFG)])
def test_ipo_short_trajectory(golden, tag, cfg):
    g, geo = golden("ipo"), golden("geom")
    trace = []
    zo.ipo_fit(g[f"{tag}_x0"], geo["db_2d"][:, :, :2], geo["K"], cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"],
               cfg["IPO_minScaleT"], cfg["IPO_maxScaleT"], iters=10, trace=trace)
    assert rel_err(trace[9][0], g[f"{tag}_q_traj"][9]) < 2e-5
    assert rel_err(trace[9][1], g[f"{tag}_s_traj"][9]) < 2e-5
    assert rel_err(zo.quaternion_to_matrix(g[f"{tag}_q_final"]), g[f"{tag}_R_final"]) < 1e-6


def test_oil_teacher_forced_step(golden):
    g, geo = golden("tf500"), golden("geom")
    W = zo.make_weights(seed=0)
    grad, T = zo.gradient_field(geo["db_2d"][:, :, :2], g["x_in"], geo["K"], conf=geo["db_2d"][:, :, 2].copy())
    _, out = zo.pc_sampler_step(W, (g["x_in"] + grad).astype(np.float32), g["t"])
    assert rel_err(out, g["x_out"]) < 1e-5 and rel_err(T, g["T_out"]) < 1e-5


def test_oil_loop_first_steps(golden):
    """Cumulative parity over the first 100 steps (the full 1000-step loop is pinned by
    oracle/gen_golden.py; here the run is kept short for the CPU suite)."""
    g, geo = golden("oil"), golden("geom")
    W = zo.make_weights(seed=0)
    x_rot = np.einsum("bij,bnj->bni", g["R"], g["x0"]).astype(np.float32)
    ts = zo.oil_time_grid()[:100]
    _, _, d = zo.oil_loop_schedule(W, x_rot, g["T"], geo["db_2d"][:, :, :2], geo["K"], geo["db_2d"][:, :, 2].copy(),
                                   ts, 200, dump_steps=(0, 9, 99))
    steps = list(g["steps"])
    for s, tol in ((0, 5e-6), (9, 5e-5), (99, 5e-4)):
        assert rel_err(d[s], g["poses"][steps.index(s)]) < tol


def test_damped_loop_golden_is_self_consistent(golden):
    """tests/golden/oil_small.npz (reference run with post_dense x 0.05): stored poses and MPJPE agree."""
    gs = golden("oil_small")
    gt = zo.make_synthetic_dataset(16, seed=7, n_clusters=3)["db_3d"].astype(np.float64)
    m = np.array([zo.mpjpe(gs["x_final"][n], gt[n]) for n in range(16)])
    assert np.abs(m - gs["mpjpe"]).max() < 1e-12 and 0.3 < m.mean() < 1.0


def test_procrustes_and_eval_multi(golden):
    g = golden("eval")
    preds, gts = g["preds"], g["gts"]
    N, S = preds.shape[:2]
    al = np.stack([zo.procrustes_align(preds[n, s], gts[n]) for n in range(N) for s in range(S)])
    assert rel_err(al, g["aligned"]) < 1e-9
    for p2 in (0, 1):
        agg, res, idx = zo.eval_multi(preds, gts, protocol2=bool(p2), actions=g["actions"])
        assert abs(agg - float(g[f"agg_p{p2}"])) < 1e-12
        assert np.abs(res - g[f"min_p{p2}"]).max() < 1e-12
        assert np.array_equal(idx, g[f"idx_p{p2}"])  # selection indices bit-exact
    assert g["idx_p0"][5] == 0  # exact tie -> first index wins
    agg, _, _ = zo.eval_multi(preds, gts, protocol2=True)
    assert abs(agg - float(g["agg_pw3d_p1"])) < 1e-12
    min_pred = preds[np.arange(N), g["idx_p0"]]
    assert abs(zo.compute_pck(gts, min_pred) - float(g["pck"])) < 1e-12
    assert abs(zo.compute_auc(gts, min_pred) - float(g["auc"])) < 1e-12


def test_c1_golden_matches_the_synthetic_generator(golden):
    """tests/golden/c1.npz stores only the reference's outputs; its inputs are regenerated from the seed.  The
    stored per-pose MPJPE must be reproducible from the stored final poses and the regenerated ground truth."""
    g = golden("c1")
    ds = zo.make_synthetic_dataset(1024, seed=int(g["seed"]), n_clusters=1)
    gt = ds["db_3d"].astype(np.float64)
    m = np.array([zo.mpjpe(g["x_final"][n], gt[n]) for n in range(1024)])
    assert np.abs(m - g["mpjpe"]).max() < 1e-12
    assert g["R"].shape == (1024, 3, 3) and np.abs(np.linalg.det(g["R"].astype(np.float64)) - 1).max() < 1e-5
    # first step of the loop from the reference's IPO output: finite and at pose scale
    x0 = zo.init_hypothesis(ds["clusters"], 0, 1024)
    x_rot = np.einsum("bij,bnj->bni", g["R"], x0).astype(np.float32)
    gr, _ = zo.gradient_field(ds["db_2d"][:8, :, :2], x_rot[:8], ds["camera_param"][:8], t=g["T"][:8].reshape(8, 1, 3),
                              conf=ds["db_2d"][:8, :, 2].copy())
    assert np.isfinite(gr).all() and np.abs(gr).max() < 5.0


def test_noise_bearing_updates_vp_ve(golden):
    """Ancestral sampling (VP / VE), Langevin corrector (2 steps, batch-mean norms) and annealed Langevin dynamics with
    injected noise, recorded from the reference's classes (sampling.py:208-324) on VPSDE / VESDE."""
    W = zo.make_weights(seed=0)
    g = golden("noise_vp")
    x, t, zs = g["x"], np.float32(g["t"]), [g["z0"], g["z1"]]
    for tag, got in (("anc_vp", zo.ancestral_update_vp(W, x, t, zs[0])), ("anc_ve", zo.ancestral_update_ve(W, x, t, zs[0])),
                     ("lang", zo.langevin_update_vp(W, x, t, zs, snr=0.16, n_steps=2)),
                     ("ald", zo.langevin_update_vp(W, x, t, zs, snr=0.16, n_steps=1, ald=True))):
        assert rel_err(got[0], g[f"{tag}_x"]) < 5e-6 and rel_err(got[1], g[f"{tag}_mean"]) < 5e-6, tag


def test_kmeans_restatement_recovers_planted_clusters():
    """The numpy k-means the CUDA cluster generator is checked against (oracle/zedo_oracle.py: kmeans_lloyd)."""
    rng = np.random.default_rng(0)
    modes = rng.normal(0, 1.0, (5, 12)).astype(np.float32)
    lab_true = rng.integers(0, 5, 800)
    x = (modes[lab_true] + rng.normal(0, 0.01, (800, 12))).astype(np.float32)
    first = [int(np.flatnonzero(lab_true == k)[0]) for k in range(5)]
    c, lab, dist = zo.kmeans_lloyd(x, x[first], 5)
    assert np.array_equal(lab, lab_true) and np.abs(c - modes).max() < 5e-3 and dist.max() < 12 * 0.05 ** 2
