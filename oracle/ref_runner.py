"""The reference's OWN PyTorch path, driven on any device ("cpu" or "cuda").

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference /
``cpu_baseline`` legs may import this module; nothing under ``zedo_release_b200/`` does.

It imports the UNMODIFIED reference modules staged under ``oracle/_ref/`` by ``oracle/fetch_ref.py``
(or ``/root/reference`` itself when that tree is mounted) and replays the driver body of
``run/opt_main.py:166-222`` literally -- the same library calls in the same order, the same per-step
host round trip of ``pc_sampler`` (``sampling.py:515,525``) -- with the one change SURVEY appendix D
describes: ``device`` is a variable instead of the hard-coded ``torch.device("cuda")``
(``run/opt_main.py:66``).  ``run/opt_main_infant.py`` does not import as shipped and is not replayed here.

Used for three things:
  * parity: the CUDA kernels against the reference's own eager-PyTorch run ON THE SAME B200
    (tests/test_gpu_vs_reference.py);
  * ``bench.py --impl reference``: the reference on the box's host cores (all threads);
  * the "reference eager fp32 on a B200" row of BASELINE.md (TF32 off = torch default).
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
STAGED = os.path.join(HERE, "_ref")


def reference_root() -> str | None:
    """The staged copy if present, else a mounted reference tree, else None."""
    if os.path.isdir(os.path.join(STAGED, "lib")):
        return STAGED
    env = os.environ.get("ZEDO_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(env, "lib")):
        return env
    return None


def available() -> bool:
    return reference_root() is not None


@contextlib.contextmanager
def _isolated_lib_namespace():
    """The reference's top-level package is called ``lib`` -- so is the mirror once installed.  Import the
    reference with whatever ``lib*`` modules are registered set aside, and put them back afterwards; the
    reference's module objects stay alive through the namespace ``load()`` returns."""
    saved = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.")}
    for k in saved:
        del sys.modules[k]
    try:
        yield
    finally:
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            del sys.modules[k]
        sys.modules.update(saved)


_CACHE = {}


def load(root: str | None = None):
    """Import the reference's hot-path modules; returns a namespace of module objects / classes."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference sources not found: run `python oracle/fetch_ref.py` in the build container")
    if root in _CACHE:
        return _CACHE[root]
    if SHIMS not in sys.path:
        sys.path.append(SHIMS)  # behind site-packages: a real ml_collections / prettytable wins
    import warnings
    try:  # the reference's model.py imports torchvision.utils; importing it first keeps torchvision's own import-time
        import torchvision.utils  # noqa: F401  source inspection away from the reference's namespace package `lib`
    except Exception:
        pass
    with _isolated_lib_namespace(), warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)  # a docstring of the reference's transforms.py holds "\m"
        sys.path.insert(0, root)
        try:
            import torch
            from lib.algorithms.advanced import sde_lib, sampling, utils as mutils
            from lib.algorithms.advanced.model import ScoreModelFC_Adv
            from lib.algorithms.advanced.control_model import Control_ScoreModelFC_Adv
            from lib.algorithms.advanced import simple_zeroshot_opt as szo
            from lib.utils import transforms
            from lib.dataset.h36m import H36MDataset3D
            from lib.dataset.pw3d import PW3D
        finally:
            sys.path.remove(root)
    R = types.SimpleNamespace(root=root, torch=torch, sde_lib=sde_lib, sampling=sampling, mutils=mutils,
                              ScoreModelFC_Adv=ScoreModelFC_Adv, Control=Control_ScoreModelFC_Adv, szo=szo,
                              transforms=transforms, H36M=H36MDataset3D, PW3D=PW3D)
    _CACHE[root] = R
    return R


def ns(**kw):
    return types.SimpleNamespace(**kw)


def ref_config(device="cpu", zedo: dict | None = None):
    """The attributes the hot path reads, values of configs/optim/concat_pose_optimization_h36m.py."""
    cfg = ns(
        training=ns(sde="subvpsde", continuous=True, cond_pose_mask_prob=0.0, cond_part_mask_prob=0.0,
                    cond_joint_mask_prob=0.0),
        sampling=ns(method="pc", predictor="euler_maruyama", corrector="none", snr=0.16, n_steps_each=1,
                    probability_flow=True, noise_removal=True),
        model=ns(embedding_type="positional", scale_by_sigma=False, sigma_max=50, sigma_min=0.01,
                 num_scales=1000, beta_min=0.1, beta_max=20.0, t=0.1, ema_rate=0.9999),
        device=device,
    )
    if zedo is not None:
        cfg.ZeDO = ns(**zedo)
    return cfg


def build_model(R, W: dict, device="cpu", n_joints=17, control=False, fourier=False):
    """``ScoreModelFC_Adv(config, 17, 3, 1024, 512, cond_dim=3)`` (run/opt_main.py:69-77) with the given state dict."""
    torch = R.torch
    cfg = ref_config(device)
    if fourier:
        cfg.model.embedding_type = "fourier"
    cls = R.Control if control else R.ScoreModelFC_Adv
    model = cls(cfg, n_joints=n_joints, joint_dim=3, hidden_dim=1024, embed_dim=512, cond_dim=3)
    missing, unexpected = model.load_state_dict({k: torch.tensor(np.asarray(v)) for k, v in W.items()}, strict=False)
    assert set(missing) <= {"sigmas"}, missing
    assert not unexpected, unexpected
    model.to(device)
    model.eval()
    return model


def make_sampling_fn(R, device, batch, n_joints=17, sampling_eps=0.01):
    """run/opt_main.py:141-160."""
    cfg = ref_config(device)
    sde = R.sde_lib.subVPSDE(beta_min=cfg.model.beta_min, beta_max=cfg.model.beta_max, N=cfg.model.num_scales,
                             T=cfg.model.t)
    cfg.sampling.probability_flow = True
    fn = R.sampling.get_sampling_fn(cfg, sde, (batch, n_joints, 3), lambda x: x, sampling_eps, device=device)
    return sde, fn


def _sync(torch, device):
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()


def ipo(R, denoise_x, condition, K, zedo, device, iters=None, pelvis=(0, 0)):
    """run/opt_main.py:175-196 (``pelvis=(0, 3)``: the SyRIP pelvis of run/opt_main_infant.py:258-261): returns
    (rot_mat, T)."""
    torch = R.torch
    optim = torch.optim
    if pelvis[0] != pelvis[1]:
        p2 = (condition[:, pelvis[0], :] + condition[:, pelvis[1], :]) / 2
    else:
        p2 = condition[:, pelvis[0], :]
    pelvis_keypoints = torch.cat((p2, torch.ones((condition.shape[0], 1), device=device)), axis=-1)
    T = torch.inverse(K).bmm(pelvis_keypoints[:, :, None]).permute(0, 2, 1)
    T = T / torch.norm(T, dim=-1, keepdim=True) * zedo["IPO_T"]
    rot_opt = R.szo.RotOpt(denoise_x.shape[0], axis=zedo["RotAxes"], minT=zedo["IPO_minScaleT"],
                           maxT=zedo["IPO_maxScaleT"])
    rot_opt.to(device)
    rot_optimizer = optim.Adam(rot_opt.parameters(), lr=0.1)
    criterion = torch.nn.L1Loss(reduction='none')
    keypoint_list = list(zedo["IPO_keylist"])
    for _ in range(zedo["IPO_iterations"] if iters is None else iters):
        rot_optimizer.zero_grad()
        rot2d = rot_opt(denoise_x[:, keypoint_list, :], T, K)
        loss = criterion(rot2d[:, :, :2], condition[:, keypoint_list, :2])
        loss = torch.mean(loss)
        loss.backward()
        rot_optimizer.step()
    with torch.no_grad():
        T = T * torch.clamp(rot_opt.scale, min=zedo["IPO_minScaleT"], max=zedo["IPO_maxScaleT"])
        rot_mat = rot_opt.generate_matrix()
    return rot_mat.detach(), T.detach()


def oil(R, model, sampling_fn, sde, denoise_x, T, condition, K, conf, device, steps=1000, sampling_eps=0.01,
        phase_switch=None, dump_steps=(), n_run=None):
    """run/opt_main.py:197-222, literally (per-step numpy round trip included).  ``n_run`` < steps runs only the
    first n_run steps of the ``steps``-long schedule (bounded samples of the bench).  Returns (results np [B,J,3],
    T, {step: np state after that step})."""
    torch = R.torch
    sample_num = steps
    timestamp = torch.linspace(sde.T, sampling_eps, sample_num, device=device)
    switch = sample_num // 5 if phase_switch is None else phase_switch
    dumps = {}
    results = denoise_x.detach().cpu().numpy()
    with torch.no_grad():
        for i in range(0, sample_num if n_run is None else n_run):
            if i < switch:
                joint_gradient = R.szo.gradient_field_gen(condition, denoise_x, K, t=T, conf=conf, returnT=False)
            else:
                joint_gradient, T = R.szo.gradient_field_gen(condition, denoise_x, K, conf=conf, returnT=True)
            denoise_x += joint_gradient
            trajs, results = sampling_fn(model, condition=condition * 0, gradient=joint_gradient,
                                         denoise_x=denoise_x, t=timestamp[i], t_step=i, args=None)
            denoise_x = torch.tensor(results).to(device)
            if i in dump_steps:
                dumps[i] = results.copy()
    return results, T, dumps


def run_pipeline(R, model, db_2d, K_np, clusters, zedo, device="cpu", hypo=1, steps=None, ipo_iters=None,
                 n_run=None, fixed_RT=None, dump_steps=(), use_conf=True, phase_switch=None):
    """The body of ``for sid in range(args.hypo)`` (run/opt_main.py:166-222) -> (batch_results [B,S,J,3] np,
    info dict with the IPO outputs of the last hypothesis and wall-clock seconds per phase)."""
    torch = R.torch
    B, J = db_2d.shape[0], db_2d.shape[1]
    sde, sampling_fn = make_sampling_fn(R, device, B, J, zedo["sampling_eps"])
    sample_poses = np.asarray(clusters, dtype=np.float32)
    batch_results = []
    info = {"t_ipo": 0.0, "t_oil": 0.0}
    for sid in range(hypo):
        noisy = torch.ones((B, J, 3)) * torch.tensor(sample_poses - sample_poses[:, 0:1, :])[sid:sid + 1, :, :]
        condition = torch.tensor(db_2d[:, :, :2], device=device).float()
        conf = torch.tensor(db_2d[:, :, 2], device=device).float() if use_conf else None
        denoise_x = noisy[:].clone().to(device)
        K = torch.tensor(K_np, device=device).float()
        _sync(torch, device)
        t0 = time.perf_counter()
        if fixed_RT is None:
            rot_mat, T = ipo(R, denoise_x, condition, K, zedo, device, iters=ipo_iters)
        else:
            rot_mat = torch.tensor(fixed_RT[0], device=device)
            T = torch.tensor(fixed_RT[1], device=device).reshape(B, 1, 3)
        _sync(torch, device)
        info["t_ipo"] += time.perf_counter() - t0
        info["R"], info["T0"] = rot_mat.cpu().numpy(), T.cpu().numpy()
        t0 = time.perf_counter()
        with torch.no_grad():
            denoise_x = rot_mat.bmm(denoise_x.permute(0, 2, 1)).permute(0, 2, 1).contiguous()
        info["x_rot"] = denoise_x.cpu().numpy()
        results, T, dumps = oil(R, model, sampling_fn, sde, denoise_x, T, condition, K, conf, device,
                                steps=zedo["OIL_iterations"] if steps is None else steps,
                                sampling_eps=zedo["sampling_eps"], phase_switch=phase_switch, dump_steps=dump_steps,
                                n_run=n_run)
        _sync(torch, device)
        info["t_oil"] += time.perf_counter() - t0
        info["T"], info["dumps"] = T.cpu().numpy(), dumps
        batch_results.append(results)
    return np.swapaxes(np.array(batch_results), 0, 1), info
