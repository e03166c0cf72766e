"""The reference's own driver file, ``run/opt_main.py`` (staged unmodified under oracle/_ref/ by
oracle/fetch_ref.py), executed end to end three ways in the same synthetic working directory:

  1. against the reference's own ``lib`` (its eager PyTorch path on this GPU),
  2. through ``python -m zedo_release_b200.dropin`` (mirror installed, hot path on the sm_100a kernels,
     dataset loaders / everything else from the reference checkout),
  3. with only PYTHONPATH + ZEDO_REFERENCE_ROOT changed.

The MPJPE tables the driver prints (eval_multi, protocol 1 and 2, per action + average) must agree.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from dropin_workdir import make_workdir, make_workdir_3dhp, make_workdir_pw3d, parse_3dhp, parse_means, parse_table

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
SHIMS = os.path.join(ROOT, "oracle", "shims")


def _run(cmd, cwd, pythonpath, extra_env=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(pythonpath)
    env.pop("ZEDO_REFERENCE_ROOT", None)
    env.update(extra_env or {})
    p = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, f"{' '.join(cmd)} failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return p.stdout


@pytest.fixture(scope="module")
def workdir(built_lib, tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: there is no CPU fallback")
    if not os.path.isfile(os.path.join(REF, "run", "opt_main.py")):
        pytest.fail("oracle/_ref is not staged: run `python oracle/fetch_ref.py` in the build container before gpurun")
    path = str(tmp_path_factory.mktemp("zedo_workdir"))
    return path, make_workdir(path, n_poses=60, hypo=2, ipo=10, oil=100)


def _driver_args(w):
    return ["--config", w["config"], "--ckpt_dir", w["ckpt_dir"], "--ckpt_name", w["ckpt_name"], "--hypo", "2"]


def test_unmodified_driver_reference_vs_mirror(workdir):
    path, w = workdir
    driver = os.path.join(REF, "run", "opt_main.py")
    # 1. the reference itself
    out_ref = _run([sys.executable, driver] + _driver_args(w), path, [REF, SHIMS])
    # 2. the launcher: mirror installed over the checkout
    out_mir = _run([sys.executable, "-m", "zedo_release_b200.dropin", REF, "run/opt_main.py"] + _driver_args(w), path,
                   [ROOT, SHIMS])
    # 3. path order only (the mirror's `lib` shadows the reference's; the rest falls through)
    out_pp = _run([sys.executable, driver] + _driver_args(w), path,
                  [ROOT, os.path.join(ROOT, "zedo_release_b200"), SHIMS], {"ZEDO_REFERENCE_ROOT": REF})
    t_ref, t_mir, t_pp = parse_table(out_ref), parse_table(out_mir), parse_table(out_pp)
    assert set(t_ref) == {"p1", "p2"} and set(t_mir) == {"p1", "p2"}, (out_ref[-2000:], out_mir[-2000:])
    for proto in ("p1", "p2"):
        a, b, c = np.array(t_ref[proto]), np.array(t_mir[proto]), np.array(t_pp[proto])
        assert np.isfinite(a).all() and a[-1] > 0.01  # a real error level in metres, not a degenerate run
        # printed with 5 decimals (1e-5 m = 0.01 mm): per action and average within 0.1 mm of the reference's
        assert np.abs(a - b).max() <= 1e-4, (proto, a, b)
        assert np.array_equal(b, c), (proto, b, c)   # both routes into the mirror are the same code
    print("reference:", t_ref, "\nmirror:   ", t_mir)


def test_detected_2d_and_gt_flag(workdir):
    """--gt switches the 2D input (run/opt_main.py:45,85); conf comes from the detection file otherwise."""
    path, w = workdir
    driver = os.path.join(REF, "run", "opt_main.py")
    out_ref = _run([sys.executable, driver] + _driver_args(w) + ["--gt"], path, [REF, SHIMS])
    out_mir = _run([sys.executable, "-m", "zedo_release_b200.dropin", REF, "run/opt_main.py"] + _driver_args(w) +
                   ["--gt"], path, [ROOT, SHIMS])
    t_ref, t_mir = parse_table(out_ref), parse_table(out_mir)
    for proto in ("p1", "p2"):
        assert np.abs(np.array(t_ref[proto]) - np.array(t_mir[proto])).max() <= 1e-4


def test_unmodified_inference_driver_3dpw_format(built_lib, tmp_path_factory):
    """``run/inference.py`` (the in-the-wild driver, north_star: "run/opt_main.py and run/inference.py drive it
    unchanged") executed as a file with the shipped 3DPW config on a synthetic ``data/3dpw/pw3d_test.npz``: the
    reference's own ``lib``, then the same file over the mirror.  ``PW3D`` (the reference's loader, its own
    ``eval_multi``) is taken from the checkout both times; sampler, score network, ``gradient_field_gen`` and ``RotOpt``
    are the sm_100a kernels the second time.  The printed MPJPE / PA-MPJPE and the saved ``results.npy`` must agree."""
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: there is no CPU fallback")
    if not os.path.isfile(os.path.join(REF, "run", "inference.py")):
        pytest.fail("oracle/_ref is not staged: run `python oracle/fetch_ref.py` in the build container before gpurun")
    path = str(tmp_path_factory.mktemp("zedo_workdir_pw3d"))
    w = make_workdir_pw3d(path, n_poses=48, hypo=2, ipo=10, oil=100)
    args = _driver_args(w) + ["--eval", "--gt"]
    out_ref = _run([sys.executable, os.path.join(REF, "run", "inference.py")] + args, path, [REF, SHIMS])
    res_ref = np.load(os.path.join(path, "results.npy"))
    os.remove(os.path.join(path, "results.npy"))
    out_mir = _run([sys.executable, "-m", "zedo_release_b200.dropin", REF, "run/inference.py"] + args, path, [ROOT, SHIMS])
    res_mir = np.load(os.path.join(path, "results.npy"))
    m_ref, m_mir = parse_means(out_ref), parse_means(out_mir)
    assert set(m_ref) == {"p1", "p2"} and set(m_mir) == {"p1", "p2"}, (out_ref[-2000:], out_mir[-2000:])
    assert res_ref.shape == res_mir.shape == (48, 2, 17, 3) and np.isfinite(res_mir).all()
    for proto in ("p1", "p2"):
        assert m_ref[proto] > 0.01 and abs(m_ref[proto] - m_mir[proto]) <= 1e-4, (proto, m_ref, m_mir)  # 0.1 mm
    # per-pose agreement of the saved hypotheses (metres): 10 Adam iterations + 100 OIL steps stay far from chaos
    assert np.abs(res_ref - res_mir).max() < 2e-3 and np.abs(res_ref - res_mir).mean() < 1e-4
    print("reference:", m_ref, "mirror:", m_mir, "max |d results|:", float(np.abs(res_ref - res_mir).max()))


def test_unmodified_driver_3dhp_format_pck_auc(built_lib, tmp_path_factory):
    """``run/opt_main.py`` with the shipped MPI-INF-3DHP config on a synthetic ``data/3dhp/mpii3d_test.pkl``: the
    reference's ``MPII3DHP`` loader and its ``eval_multi`` (PCK, AUC, hypothesis std, per-action table;
    lib/dataset/mpii3dHP.py:424-511) come from the checkout both times; over the mirror its ``compute_PCK`` /
    ``compute_AUC`` / ``align_to_gt`` imports resolve to the device kernels (zedo_pck_counts, the Procrustes kernel)."""
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: there is no CPU fallback")
    path = str(tmp_path_factory.mktemp("zedo_workdir_3dhp"))
    w = make_workdir_3dhp(path, n_poses=56, hypo=2, ipo=10, oil=100)
    args = _driver_args(w) + ["--gt"]
    driver = os.path.join(REF, "run", "opt_main.py")
    out_ref = _run([sys.executable, driver] + args, path, [REF, SHIMS])
    out_mir = _run([sys.executable, "-m", "zedo_release_b200.dropin", REF, "run/opt_main.py"] + args, path, [ROOT, SHIMS])
    r_ref, r_mir = parse_3dhp(out_ref), parse_3dhp(out_mir)
    assert set(r_ref) == {"p1", "p2"} and set(r_mir) == {"p1", "p2"}, (out_ref[-3000:], out_mir[-3000:])
    for proto in ("p1", "p2"):
        a, b = r_ref[proto], r_mir[proto]
        assert np.isfinite(a["row"]).all() and a["row"][-1] > 0.01
        assert np.abs(np.array(a["row"]) - np.array(b["row"])).max() <= 1e-4          # 0.1 mm, per action and average
        # 56 x 17 joints: one joint crossing a 5 mm threshold moves PCK by 0.105 and AUC by 0.0034 per cent
        assert abs(a["pck"] - b["pck"]) <= 0.22 and abs(a["auc"] - b["auc"]) <= 0.05, (a, b)
        assert np.abs(np.array(a["std"]) - np.array(b["std"])).max() <= 1e-4
    print("reference:", r_ref, "\nmirror:   ", r_mir)
