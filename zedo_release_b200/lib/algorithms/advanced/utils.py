"""Model/score helpers with the reference's interface (lib/algorithms/advanced/utils.py)."""
import numpy as np
import torch

from . import sde_lib

_MODELS = {}


def register_model(cls=None, *, name=None):
    """Decorator registering a model class; a duplicate name raises ValueError (utils.py:633-648)."""
    def _register(c):
        key = c.__name__ if name is None else name
        if key in _MODELS:
            raise ValueError(f'Already registered model with name: {key}')
        _MODELS[key] = c
        return c
    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]


def get_sigmas(config):
    """SMLD noise levels (utils.py:655-666)."""
    return np.exp(np.linspace(np.log(config.model.sigma_max), np.log(config.model.sigma_min),
                              config.model.num_scales))


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """Real-part-first quaternions (..., 4) -> rotation matrices (..., 3, 3); the quaternion is NOT
    normalised by the caller, two_s = 2 / |q|^2 (utils.py:59-88).  Interface-parity helper for
    ``RotOpt.generate_matrix``; the IPO kernel (csrc/ipo.cu) holds its own copy."""
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def get_model_fn(model, train=False):
    """model_fn(x, labels, condition, mask); calls model.eval()/train() on every call like the
    reference (utils.py:703-730)."""
    def model_fn(x, labels, condition, mask):
        model.train() if train else model.eval()
        return model(x, labels, condition, mask)
    return model_fn


def get_score_fn(sde, model, train=False, continuous=False):
    """Score function wrapper (utils.py:734-800).  VP / sub-VP: labels = 999 t (continuous or sub-VP)
    and score = -model/std; VE: labels = marginal std (continuous) or the rounded level index."""
    model_fn = get_model_fn(model, train=train)
    if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
        def score_fn(x, t, condition, mask):
            if continuous or isinstance(sde, sde_lib.subVPSDE):
                out = model_fn(x, t * 999, condition, mask)
                std = sde.marginal_prob(torch.zeros_like(x), t)[1]
            else:
                labels = t * (sde.N - 1)
                out = model_fn(x, labels, condition, mask)
                std = sde.sqrt_1m_alphas_cumprod.to(labels.device)[labels.squeeze(-1).long()]
            return -out / std[:, None, None]
    elif isinstance(sde, sde_lib.VESDE):
        def score_fn(x, t, condition, mask):
            if continuous:
                labels = sde.marginal_prob(torch.zeros_like(x), t)[1]
            else:
                labels = torch.round((sde.T - t) * (sde.N - 1)).long()
            return model_fn(x, labels, condition, mask)
    else:
        raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
    # what the fused update kernels need to know about this closure (sampling.py of the mirror)
    score_fn.zedo = dict(sde=sde, model=model, continuous=continuous, train=train)
    return score_fn


def to_flattened_numpy(x):
    return x.detach().cpu().numpy().reshape((-1,))


def from_flattened_numpy(x, shape):
    return torch.from_numpy(x.reshape(shape))


def _pck_percentages(gts, preds, eval_joints):
    """PCK (per cent) at the 31 thresholds linspace(0, 150, 31) mm, counted by ``zedo_pck_counts`` (csrc/eval.cu)."""
    from zedo_release_b200 import engine
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.as_tensor(np.ascontiguousarray(np.asarray(gts, dtype=np.float64)), device=dev)
    p = torch.as_tensor(np.ascontiguousarray(np.asarray(preds, dtype=np.float32)), device=dev)[:, None]
    return engine.pck_curve(p, g, joint_subset=None if eval_joints is None else [int(j) for j in eval_joints])


def compute_PCK(gts, preds, scales=1000, eval_joints=None, threshold=150):
    """MPI-INF-3DHP PCK with the reference's signature (utils.py:814-836; ``scales`` is ignored there too: the error is
    always taken in millimetres): per cent of the joints whose error is strictly below ``threshold`` mm.  gts / preds
    [N, J, 3] in metres.  Thresholds on the 5 mm grid of the AUC (0, 5, ..., 150) are counted on the device."""
    k = float(threshold) / 5.0
    if not (0 <= k <= 30 and k == int(k)):
        raise NotImplementedError("compute_PCK: thresholds on the 0, 5, ..., 150 mm grid of compute_AUC are served")
    return float(_pck_percentages(gts, preds, eval_joints)[int(k)])


def compute_AUC(gts, preds, scales=1000, eval_joints=None):
    """Mean of the PCK over thresholds linspace(0, 150, 31) mm (utils.py:839-849, mimicking mpii_compute_3d_pck.m)."""
    return float(np.mean(_pck_percentages(gts, preds, eval_joints)))
