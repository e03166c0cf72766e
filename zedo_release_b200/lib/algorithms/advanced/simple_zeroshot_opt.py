"""``gradient_field_gen`` and ``RotOpt`` with the reference's interface
(lib/algorithms/advanced/simple_zeroshot_opt.py), backed by csrc/geom.cu and csrc/ipo.cu."""
import torch
import torch.nn as nn

from zedo_release_b200 import engine
from .utils import quaternion_to_matrix


class _RotOptProject(torch.autograd.Function):
    """uv = proj(K (R(q) x + T clamp(scale))) with the analytic backward of csrc/ipo.cu, so the
    reference's ``loss.backward(); Adam.step()`` loop (run/opt_main.py:186-193) runs on two kernels
    per iteration instead of ~60."""

    @staticmethod
    def forward(ctx, q, scale, xk, T0, K, minT, maxT):
        ctx.save_for_backward(q, scale, xk, T0, K)
        ctx.lim = (minT, maxT)
        return engine.rotopt_forward(q, scale, xk, T0, K, minT, maxT)

    @staticmethod
    def backward(ctx, d_uv):
        q, scale, xk, T0, K = ctx.saved_tensors
        d_q, d_s = engine.rotopt_backward(q, scale, xk, T0, K, ctx.lim[0], ctx.lim[1], d_uv.contiguous())
        return d_q, d_s, None, None, None, None, None


class RotOpt(nn.Module):
    """Per-pose rotation (unnormalised quaternion, only the axes in ``axis`` trainable) and scale
    (simple_zeroshot_opt.py:8-31).  Parameter names and shapes match the reference."""

    def __init__(self, batch_size=100, axis='y', minT=0.5, maxT=2):
        super().__init__()
        self.rot_vect = nn.Parameter(torch.ones((batch_size, 1)))
        for axe in axis:
            setattr(self, 'rot_vect_%s' % axe, nn.Parameter(torch.zeros((batch_size, 1))))
        self.identity = nn.Parameter(torch.eye(3), requires_grad=False)
        self.batch_size = batch_size
        self.scale = nn.Parameter(torch.ones((batch_size, 1, 1)))
        self.minT, self.maxT = minT, maxT

    def _quaternion(self):
        zeros = torch.zeros((self.batch_size, 1), device=self.rot_vect.device)
        return torch.cat([self.rot_vect, getattr(self, 'rot_vect_x', zeros), getattr(self, 'rot_vect_y', zeros),
                          getattr(self, 'rot_vect_z', zeros)], dim=-1)

    def forward(self, x, T, K):
        """x [B,k,3] key joints, T [B,1,3], K [B,3,3] -> projected 2D [B,k,2]."""
        B = x.shape[0]
        return _RotOptProject.apply(self._quaternion(), self.scale.reshape(B), x.contiguous().float(),
                                    T.reshape(B, 3).contiguous().float(), K.contiguous().float(),
                                    float(self.minT), float(self.maxT))

    def generate_matrix(self):
        return quaternion_to_matrix(self._quaternion())


def perpendicular_distance(point, vector):
    """(point . vector) vector - point (simple_zeroshot_opt.py:33-36); interface parity."""
    return torch.sum(point * vector, dim=-1, keepdim=True) * vector - point


def gradient_field_gen(key2d, key3d, K, noise_type=None, t=None, conf=None, returnT=False, norm_true=None,
                       previous_T=None):
    """Gradient of the 3D key points towards their camera rays (simple_zeroshot_opt.py:46-125).

    key2d [b,n,2], key3d [b,n,3], K [b,3,3], conf [b,n] (clamped IN PLACE to [1e-4, 1] like the
    reference) or None, t [b,1,3] fixed translation or None (least-squares solve, sign flip on
    T_z < 0).  Returns gradient, or (gradient, T) when returnT.
    """
    std = 0.0001
    if conf is not None and not (conf.is_cuda and conf.dtype == torch.float32 and conf.is_contiguous()):
        raise ValueError("conf must be a contiguous float32 CUDA tensor (it is clamped in place)")
    gradient, T = engine.grad_field(key2d, key3d, K, conf=conf, T=t)
    if t is not None:
        T = t
    if noise_type == 'gaussian':
        gradient = gradient + std * torch.randn(*gradient.shape).to(gradient.device) * t
    elif noise_type == 'uniform':
        gradient = gradient + std * (torch.randn(*gradient.shape) - 0.5).to(gradient.device)
    return (gradient, T) if returnT else gradient
