"""``ScoreModelFC_Adv`` with the reference's constructor, parameter names and call signature
(lib/algorithms/advanced/model.py:97-298); ``forward`` runs the fused tcgen05 layer kernels.

The module only *holds* the parameters (so ``state_dict`` / ``load_state_dict`` / ``.to(device)``
and the checkpoint format of run/opt_main.py:120-137 work unchanged).  On the first forward, and
whenever a parameter changes, the weights are packed into a ``zedo_plan`` (fp16 hi/lo blocked
layout, concatenated time-projection matrix, GroupNorm tables).  There is no PyTorch fallback:
training mode, Fourier embeddings and hidden_dim != 1024 raise instead of silently diverging.
"""
import functools
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from zedo_release_b200 import engine


def get_sigmas(config):
    """SMLD noise levels (model.py:68-78)."""
    return np.exp(np.linspace(np.log(config.model.sigma_max), np.log(config.model.sigma_min),
                              config.model.num_scales))


def get_timestep_embedding(timesteps, embedding_dim, max_positions=10000):
    """Sinusoidal embedding (model.py:81-95); interface parity -- the kernels build the embedding
    themselves (csrc/mlp_simt.cu: timestep_embedding_kernel)."""
    assert len(timesteps.shape) == 1
    half_dim = embedding_dim // 2
    emb = math.log(max_positions) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float32, device=timesteps.device) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1), mode='constant')
    assert emb.shape == (timesteps.shape[0], embedding_dim)
    return emb


class GaussianFourierProjection(nn.Module):
    """Gaussian random features for encoding time steps (model.py:27-36).  The score modules hand ``W`` to the plan
    (state_dict entry ``gauss_proj.W``), which evaluates the features on the device; ``forward`` is kept for callers
    that use the projection on its own."""

    def __init__(self, embed_dim, scale=30.):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embed_dim // 2) * scale, requires_grad=False)

    def forward(self, x):
        x_proj = x[:, None] * self.W[None, :] * 2 * np.pi
        return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


class _PlanCache:
    """Packed-weights cache shared by the score modules: rebuilt when any parameter is modified
    (tensor version counters), re-created with a larger capacity when a bigger batch arrives.
    Writes through ``param.data`` (which carry their own version counter) are invisible to the key: the mirror's
    ``ExponentialMovingAverage.copy_to / restore`` write through the parameter for that reason, and code that edits
    ``.data`` by hand calls ``module.invalidate_plan()``."""

    def __init__(self):
        self.plan, self.key = None, None

    def invalidate(self):
        if self.plan is not None:
            self.plan.close()
        self.plan, self.key = None, None

    def get(self, module, batch, n_joints, hidden, embed, n_blocks):
        params = dict(module.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("ScoreModelFC_Adv.forward needs the model on a CUDA device: "
                               "zedo_release_b200 has no CPU path")
        key = (dev.index, tuple((k, v.data_ptr(), v._version) for k, v in params.items()))
        if self.plan is None or self.key != key or batch > self.plan.capacity:
            if self.plan is not None:
                self.plan.close()
            cap = max(1024, 1 << (int(batch) - 1).bit_length())
            state = {k: v.detach() for k, v in module.state_dict().items()}
            self.plan = engine.ScorePlan(state, n_joints=n_joints, hidden=hidden, embed=embed, n_blocks=n_blocks,
                                         max_batch=cap, device=dev.index if dev.index is not None else 0)
            self.key = key
        return self.plan


class ScoreModelFC_Adv(nn.Module):
    """Independent condition feature projection layers for each block (model.py:97-152)."""

    def __init__(self, config, n_joints=17, joint_dim=3, hidden_dim=64, embed_dim=32, cond_dim=2, n_blocks=2):
        super().__init__()
        self.config = config
        self.n_joints, self.joint_dim, self.n_blocks = n_joints, joint_dim, n_blocks
        self.hidden_dim, self.embed_dim = hidden_dim, embed_dim
        self.act = nn.SiLU()
        self.pre_dense = nn.Linear(n_joints * joint_dim, hidden_dim)
        self.pre_dense_t = nn.Linear(embed_dim, hidden_dim)
        self.pre_gnorm = nn.GroupNorm(32, num_channels=hidden_dim)
        self.dropout = nn.Dropout(p=0.25)
        self.time_embedding_type = config.model.embedding_type.lower()
        if self.time_embedding_type == 'fourier':
            self.gauss_proj = GaussianFourierProjection(embed_dim=embed_dim)
        elif self.time_embedding_type == 'positional':
            self.posit_proj = functools.partial(get_timestep_embedding, embedding_dim=embed_dim)
        else:
            assert 0
        self.shared_time_embed = nn.Sequential(nn.Linear(embed_dim, embed_dim), self.act)
        self.register_buffer('sigmas', torch.tensor(get_sigmas(config)))
        for idx in range(n_blocks):
            setattr(self, f'b{idx+1}_dense1', nn.Linear(hidden_dim, hidden_dim))
            setattr(self, f'b{idx+1}_dense1_t', nn.Linear(embed_dim, hidden_dim))
            setattr(self, f'b{idx+1}_gnorm1', nn.GroupNorm(32, num_channels=hidden_dim))
            setattr(self, f'b{idx+1}_dense2', nn.Linear(hidden_dim, hidden_dim))
            setattr(self, f'b{idx+1}_dense2_t', nn.Linear(embed_dim, hidden_dim))
            setattr(self, f'b{idx+1}_gnorm2', nn.GroupNorm(32, num_channels=hidden_dim))
        self.post_dense = nn.Linear(hidden_dim, n_joints * joint_dim)
        self.cond_pose_mask_prob = config.training.cond_pose_mask_prob
        self.cond_part_mask_prob = config.training.cond_part_mask_prob
        self.cond_joint_mask_prob = config.training.cond_joint_mask_prob
        self._plans = _PlanCache()
        self.gemm_mode = engine.DEFAULT_MODE

    def invalidate_plan(self):
        """Drop the packed weights; the next forward re-packs them from the current parameters (needed only after
        in-place edits through ``param.data``, which do not move the parameters' version counters)."""
        self._plans.invalidate()

    def zedo_plan(self, batch):
        """The packed plan for a batch of this size (also used by the fused sampler fast path)."""
        if self.joint_dim != 3:
            raise NotImplementedError("joint_dim must be 3")
        return self._plans.get(self, batch, self.n_joints, self.hidden_dim, self.embed_dim, self.n_blocks)

    def forward(self, batch, t, condition=None, mask=None):
        """batch [B,j,3], t [B] time labels (999 * t, all equal inside one call like every caller of
        the reference: vec_t = ones(B) * t, sampling.py:497); condition / mask are never read by the
        reference forward (model.py:225-244).  Returns [B,j,3]."""
        if self.training:
            raise NotImplementedError("zedo_release_b200 implements the inference path only (model.eval())")
        bs = batch.shape[0]
        t = torch.as_tensor(t, device=batch.device).reshape(-1)
        labels = t.unique()
        plan = self.zedo_plan(bs)
        x = batch.reshape(bs, self.n_joints, self.joint_dim)
        if labels.numel() == 1:
            res = plan.forward(x, float(labels[0]), mode=self.gemm_mode)
        else:  # per-row labels: one pass per distinct label (never happens in the shipped drivers)
            res = torch.empty_like(x, dtype=torch.float32)
            tt = t.expand(bs) if t.numel() == 1 else t
            for lab in labels:
                sel = (tt == lab).nonzero(as_tuple=True)[0]
                res[sel] = plan.forward(x[sel].contiguous(), float(lab), mode=self.gemm_mode)
        if self.config.model.scale_by_sigma:
            # used_sigmas = t for the 'fourier' embedding, sigmas[t.long()] for 'positional' (model.py:248,253)
            used_sigmas = t if self.time_embedding_type == 'fourier' else self.sigmas[t.long()]
            res = res / used_sigmas.reshape((-1, 1, 1)).to(res.dtype)
        return res
