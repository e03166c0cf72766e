"""Host-side logic that needs no GPU: sharding, schedules, synthetic generators, the mirror's
registries and error conventions, and the world_size-2 gather over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import zedo_oracle as zo
from conftest import ROOT


@pytest.fixture(scope="module")
def zr(built_lib):
    import zedo_release_b200 as zr
    return zr


def test_shard_range_matches_reference_rule(zr):
    for n, w in ((10, 3), (65536, 8), (7, 8), (1_000_000, 8), (5, 1)):
        spans = [zr.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))  # contiguous, no overlap
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)  # first n % w get +1


def test_schedule_and_synthetic_generators_match_oracle(zr, golden):
    assert np.array_equal(zr.linspace_schedule(0.1, 0.01, 1000), golden("sde")["time_grid"])
    assert np.array_equal(zr.linspace_schedule(0.1, 0.01, 1000), zo.oil_time_grid())
    from zedo_release_b200 import synthetic as sy
    a, b = zo.make_synthetic_dataset(64, seed=5, n_clusters=3), sy.make_synthetic_dataset(64, seed=5, n_clusters=3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    wa, wb = zo.make_weights(2, n_joints=12), sy.make_weights(2, n_joints=12)
    assert wa.keys() == wb.keys() and all(np.array_equal(wa[k], wb[k]) for k in wa)
    assert sy.H36M_ZEDO_CFG == zo.H36M_ZEDO_CFG and zr.axes_mask("xyz") == 7 and zr.axes_mask("z") == 4


def test_dataset_format_generators_match_oracle(zr):
    """The H36M pickle-item and 3DPW npz generators of the package equal the oracle's copies (the golden vectors of
    tests/golden/formats.npz were produced by the reference's loaders from the oracle's)."""
    from zedo_release_b200 import synthetic as sy
    ds = zo.make_synthetic_dataset(9, seed=31, n_clusters=2)
    a, b = zo.h36m_items_from_arrays(ds), sy.h36m_items_from_arrays(ds)
    for ia, ib in zip(a, b):
        assert ia["action"] == ib["action"] and ia["image_path"] == ib["image_path"]
        assert np.array_equal(ia["joint_3d_camera"], ib["joint_3d_camera"])
        assert np.array_equal(ia["joint_3d_image"], ib["joint_3d_image"])
        assert all(ia["camera_param"][k] == ib["camera_param"][k] for k in ("fx", "fy", "cx", "cy"))
    pa, pb = zo.pw3d_npz_from_arrays(ds), sy.pw3d_npz_from_arrays(ds)
    assert np.array_equal(pa["keypoints3d17_relative"], pb["keypoints3d17_relative"])
    assert np.array_equal(pa["root_cam"], pb["root_cam"])
    assert np.array_equal(pa["cam_param"].item()["f"], pb["cam_param"].item()["f"])


def test_aggregate_errors(zr):
    e = np.arange(30, dtype=np.float64)
    acts = 2 + np.arange(30) % 15
    assert zr.aggregate_errors(e) == pytest.approx(e.mean())
    assert zr.aggregate_errors(e, acts) == pytest.approx(np.mean([e[acts == a].mean() for a in range(2, 17)]))


def test_mirror_registries_and_error_conventions(zr):
    from zedo_release_b200.lib.algorithms.advanced import sampling, sde_lib, utils as mutils
    from types import SimpleNamespace as NS
    assert sampling.get_predictor("euler_maruyama") is sampling.EulerMaruyamaPredictor
    assert sampling.get_corrector("none") is sampling.NoneCorrector
    assert set(sampling._PREDICTORS) == {"euler_maruyama", "reverse_diffusion", "ancestral_sampling", "none"}
    assert set(sampling._CORRECTORS) == {"langevin", "ald", "none"}
    with pytest.raises(ValueError):  # duplicate registry name (reference sampling.py:42-43)
        sampling.register_predictor(name="none")(type("X", (), {}))
    sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20., N=1000, T=0.1)
    cfg = NS(sampling=NS(method="bogus", predictor="euler_maruyama", corrector="none", snr=0.16, n_steps_each=1,
                         probability_flow=True, noise_removal=True), training=NS(continuous=True), device="cpu")
    with pytest.raises(ValueError):  # unknown sampler (sampling.py:125)
        sampling.get_sampling_fn(cfg, sde, (4, 17, 3), lambda x: x, 0.01)
    with pytest.raises(NotImplementedError):  # ancestral sampling only supports VE/VP (sampling.py:215)
        sampling.AncestralSamplingPredictor(sde, lambda *a: None)
    with pytest.raises(NotImplementedError):
        mutils.get_score_fn(object(), None)
    # Langevin corrector + sub-VP SDE: the reference reads sde.alphas, which only VPSDE defines -> AttributeError
    with pytest.raises(AttributeError):
        sampling.LangevinCorrector(sde, lambda *a: None, 0.16, 1).update_fn(torch.zeros(2, 17, 3), torch.ones(2) * 0.05,
                                                                              None, None)
    vp = sde_lib.VPSDE(beta_min=0.1, beta_max=20., N=1000, T=1)
    xs, xm = sampling.LangevinCorrector(vp, lambda x, t, c, m: -x, 0.16, 2).update_fn(
        torch.ones(3, 17, 3), torch.ones(3) * 0.5, None, None)
    assert xs.shape == (3, 17, 3) and torch.isfinite(xs).all() and torch.isfinite(xm).all()
    rs = sde.reverse(lambda x, t, c, m: x * 0 + 2.0, probability_flow=True)  # module-level reverse-time object
    drift, diff = rs.sde(torch.ones(2, 17, 3), torch.tensor([0.1, 0.05]), None, None)
    assert rs.N == 1000 and rs.T == 0.1 and float(diff.sum()) == 0.0 and rs.beta_1 == 20.0
    b, g = zo.subvp_sde_scalars(np.float32([0.1, 0.05]))
    assert np.allclose(drift[:, 0, 0].numpy(), -0.5 * b - g ** 2 * 2.0, rtol=1e-5)
    # sub-VP schedule of the mirror == oracle scalars
    t = torch.tensor([0.1, 0.05, 0.01])
    _, g = sde.sde(torch.zeros(3, 1, 1), t)
    _, std = sde.marginal_prob(torch.zeros(3, 1, 1), t)
    assert np.allclose(g.numpy(), zo.subvp_sde_scalars(t.numpy())[1], rtol=2e-5)
    assert np.allclose(std.numpy(), zo.subvp_marginal_std(t.numpy()), rtol=1e-4)
    q = torch.tensor(np.random.default_rng(0).normal(size=(5, 4)).astype(np.float32))
    assert np.allclose(mutils.quaternion_to_matrix(q).numpy(), zo.quaternion_to_matrix(q.numpy()), atol=1e-6)


def test_mirror_model_state_dict_contract(zr):
    """Parameter names/shapes are the reference's (model.py:113-152) and a checkpoint with the
    DataParallel 'module.' prefix loads the way run/opt_main.py:125-136 does it."""
    from zedo_release_b200.lib.algorithms.advanced.model import ScoreModelFC_Adv
    from zedo_release_b200.lib.algorithms.ema import ExponentialMovingAverage
    from types import SimpleNamespace as NS
    cfg = NS(model=NS(embedding_type="positional", scale_by_sigma=False, sigma_max=50, sigma_min=0.01,
                      num_scales=1000, ema_rate=0.9999),
             training=NS(cond_pose_mask_prob=0.0, cond_part_mask_prob=0.0, cond_joint_mask_prob=0.0))
    m = ScoreModelFC_Adv(cfg, n_joints=17, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    W = zo.make_weights(seed=0)
    names = set(m.state_dict().keys())
    assert names == set(W.keys()) | {"sigmas"}
    ckpt = {"model_state_dict": {"module." + k: torch.tensor(v) for k, v in W.items()},
            "ema": ExponentialMovingAverage(m.parameters(), 0.9999).state_dict(), "step": 7}
    ckpt["model_state_dict"]["module.sigmas"] = m.sigmas.clone()
    m.load_state_dict({k[7:]: v for k, v in ckpt["model_state_dict"].items()})
    ema = ExponentialMovingAverage(m.parameters(), decay=cfg.model.ema_rate)
    ema.load_state_dict(ckpt["ema"])
    assert np.array_equal(m.b2_dense1.weight.detach().numpy(), W["b2_dense1.weight"])
    assert sum(p.numel() for p in m.parameters()) == 7_203_379
    # copy_to / restore write through the parameters: their version counters (the plan cache key) move
    v0 = m.post_dense.bias._version
    ema.store(m.parameters())
    ema.copy_to(m.parameters())
    v1 = m.post_dense.bias._version
    ema.restore(m.parameters())
    assert v0 < v1 < m.post_dense.bias._version
    assert np.array_equal(m.b2_dense1.weight.detach().numpy(), W["b2_dense1.weight"])
    with pytest.raises(RuntimeError):  # model on CPU: no CPU path
        m.eval()
        m(torch.zeros(2, 17, 3), torch.ones(2), None, None)


def _gloo_worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from zedo_release_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    lo, hi = parallel.shard_range(n, r, w)
    full = torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3)
    got = parallel.gather_rows(full[lo:hi].clone(), n)
    idx = parallel.gather_rows(torch.arange(lo, hi, dtype=torch.int32), n)
    vals = torch.arange(n, dtype=torch.float32) ** 2
    gm = parallel.global_batch_mean(vals[lo:hi])  # Langevin's batch-mean norms over uneven shards
    ok = (torch.equal(got, full) and torch.equal(idx, torch.arange(n, dtype=torch.int32))
          and abs(float(gm) - float(vals.double().mean())) < 1e-4 * float(vals.mean()))
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 64])
def test_gather_rows_world2_gloo(built_lib, tmp_path, n):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}.txt").read() for r in range(2)] == ["ok", "ok"]


def test_e4m3_emulator_known_answers():
    """tools/precision_study.py emulates the e4m3 operands of the fp8lo GEMM mode on the CPU (the study the mode was
    built on): round-to-nearest-even on 4 significant bits, min normal 2^-6, subnormal step 2^-9, saturation at 448."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("precision_study", os.path.join(ROOT, "tools", "precision_study.py"))
    ps = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ps)
    x = np.array([0.0, 1.0, -1.0, 0.3, 300.0, 500.0, -1000.0, 448.0, 1.0625, 1.1875, 2.0 ** -6, 0.001, 0.0009,
                  2.0 ** -9, 3 * 2.0 ** -10], dtype=np.float32)
    want = np.array([0.0, 1.0, -1.0, 0.3125, 288.0, 448.0, -448.0, 448.0, 1.0, 1.25, 2.0 ** -6, 2.0 ** -9, 0.0,
                     2.0 ** -9, 2.0 ** -8], dtype=np.float32)
    assert np.array_equal(ps.e4m3(x), want)
    # every value the emulator returns is on the e4m3 grid: idempotent
    r = np.random.default_rng(0).normal(0, 30, 10000).astype(np.float32)
    q = ps.e4m3(r)
    assert np.array_equal(ps.e4m3(q), q) and np.abs(q).max() <= 448
    # and the three-product split it models: hi16 + lo16 reproduces a float32 activation to ~2^-22
    a = r / 30
    hi = ps.fp16(a)
    assert np.abs(a - (hi + ps.fp16(a - hi))).max() <= 2.0 ** -21 * np.abs(a).max()


def test_mirror_falls_through_to_the_reference_checkout(built_lib):
    """zlib.install(reference_root=...): modules the mirror does not carry (dataset loaders, lib.utils.generic,
    losses) and single names it lacks (image_to_camera_frame) resolve to the reference's files, the hot-path
    modules stay the mirror's (run/opt_main.py:12-20 imports all of them).  Runs in a subprocess: the mirror is
    registered as top-level `lib` for good."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "lib")):
        pytest.skip("oracle/_ref not staged (python oracle/fetch_ref.py needs /root/reference)")
    code = (
        "import sys, warnings; warnings.simplefilter('ignore')\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.append({os.path.join(ROOT, 'oracle', 'shims')!r})\n"
        "import zedo_release_b200.lib as zlib\n"
        f"zlib.install(reference_root={ref!r})\n"
        "from lib.dataset.h36m import H36MDataset3D\n"
        "from lib.dataset.mpii3dHP import MPII3DHP\n"
        "from lib.dataset.pw3d import PW3D\n"
        "from lib.dataset.skiPose import skiPose\n"
        "from lib.dataset.custom import CustomDataset\n"
        "from lib.utils import generic\n"
        "from lib.algorithms.advanced import losses, sampling, sde_lib\n"
        "from lib.algorithms.advanced.simple_zeroshot_opt import gradient_field_gen, RotOpt\n"
        "from lib.utils.transforms import image_to_camera_frame, align_to_gt\n"
        "from lib.algorithms.advanced.utils import compute_PCK, get_score_fn, mean_cov\n"
        "import os\n"
        "f = lambda o: os.path.abspath(o.__code__.co_filename if hasattr(o, '__code__') else o.__file__)\n"
        f"ref, mir = {ref!r}, {os.path.join(ROOT, 'zedo_release_b200')!r}\n"
        "assert f(H36MDataset3D.__init__).startswith(ref) and f(generic).startswith(ref) and f(losses).startswith(ref)\n"
        "assert f(image_to_camera_frame).startswith(ref) and f(mean_cov).startswith(ref)\n"
        "assert f(sampling).startswith(mir) and f(sde_lib).startswith(mir) and f(align_to_gt).startswith(mir)\n"
        "assert f(gradient_field_gen).startswith(mir) and f(get_score_fn).startswith(mir) and f(compute_PCK).startswith(mir)\n"
        "print('ok')\n")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stderr[-2000:]


def test_reference_loaders_read_the_synthetic_working_directories(built_lib, tmp_path):
    """The working directories tests/dropin_workdir.py writes for the drop-in driver tests are read here by the
    reference's OWN loaders (lib/dataset/pw3d.py:184-227, mpii3dHP.py:255-300, h36m.py) from oracle/_ref, on the CPU: the
    arrays they build are the synthetic dataset's (3D in metres, 2D = projection, K), i.e. the files have the reference's
    formats.  Subprocess: the reference's `lib` is imported as a top-level package."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "lib")):
        pytest.skip("oracle/_ref not staged (python oracle/fetch_ref.py needs /root/reference)")
    code = (
        "import sys, os, warnings; warnings.simplefilter('ignore')\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        f"sys.path.insert(0, {ref!r}); sys.path.append({os.path.join(ROOT, 'oracle', 'shims')!r})\n"
        "import numpy as np\n"
        "from pathlib import Path\n"
        "from dropin_workdir import make_workdir, make_workdir_pw3d, make_workdir_3dhp\n"
        f"base = {str(tmp_path)!r}\n"
        "for name, mk in (('h36m', make_workdir), ('pw3d', make_workdir_pw3d), ('3dhp', make_workdir_3dhp)):\n"
        "    d = os.path.join(base, name); os.makedirs(d)\n"
        "    w = mk(d, n_poses=21, hypo=2, ipo=2, oil=5); ds = w['ds']; os.chdir(d)\n"
        "    if name == 'h36m':\n"
        "        from lib.dataset.h36m import H36MDataset3D\n"
        "        t = H36MDataset3D(Path('data', 'h36m'), 'test', gt2d=True, abs_coord=True, sample_interval=1, flip=False)\n"
        "    elif name == 'pw3d':\n"
        "        from lib.dataset.pw3d import PW3D\n"
        "        t = PW3D(Path('data', '3dpw'), 'test', gt2d=True, abs_coord=True, sample_interval=1, flip=False)\n"
        "    else:\n"
        "        from lib.dataset.mpii3dHP import MPII3DHP\n"
        "        t = MPII3DHP(Path('data', '3dhp'), 'test', gt2d=True, abs_coord=True, sample_interval=1, flip=False)\n"
        "    db3, db2, K = np.asarray(t.db_3d), np.asarray(t.db_2d), np.asarray(t.camera_param)\n"
        "    assert db3.shape == (21, 17, 3) and db2.shape[:2] == (21, 17) and K.shape == (21, 3, 3), (name, db3.shape, db2.shape)\n"
        "    want3 = ds['db_3d'] + ds['root'][:, None, :]\n"
        "    assert np.abs(db3 - want3).max() < 1e-3, (name, np.abs(db3 - want3).max())        # metres, camera frame\n"
        "    proj = np.einsum('nij,nkj->nki', ds['camera_param'].astype(np.float64), want3.astype(np.float64))\n"
        "    want2 = proj[:, :, :2] / proj[:, :, 2:3] if name == 'pw3d' else ds['db_2d'][:, :, :2]   # PW3D re-projects its 3D\n"
        "    assert np.abs(db2[:, :, :2] - want2).max() < 2e-2, (name, np.abs(db2[:, :, :2] - want2).max())   # pixels\n"
        "    assert np.abs(K - ds['camera_param']).max() < 1e-3, name\n"
        "    cl = np.load(os.path.join(d, 'clusters', ('3dhp' if name == '3dhp' else 'h36m') + '_cluster2.npy'))\n"
        "    assert cl.shape == (2, 17, 3) and cl.dtype == np.float32\n"
        "print('ok')\n")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), (p.stdout[-1500:], p.stderr[-2500:])
