"""CPU oracle for the ZeDO per-pose optimisation loop (TEST INFRASTRUCTURE ONLY).

This file is a plain numpy (float32) restatement of the reference's hot path.  It is
the checker for the CUDA kernels: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
package (``zedo_release_b200``) never imports anything from ``oracle/``.

Parity pinning: every function below is checked against the *imported reference
modules* (``/root/reference``) by ``oracle/gen_golden.py`` (run in the build container,
where the reference is mounted); that script also writes the golden vectors under
``tests/golden/`` which ``tests/test_oracle_golden.py`` replays without the reference.
The reference ships no tests and only one known-answer (the ``__main__`` demo of
``simple_zeroshot_opt.py:127-147``, first value 53.63671875) which is pinned too.

Citations ``file:line`` are relative to the reference root.
All arithmetic is float32 unless stated (the reference runs torch float32); the
evaluation functions follow the reference's numpy dtype promotion (float64 when the
ground truth is float64).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
Weights = Dict[str, np.ndarray]

# --------------------------------------------------------------------------------------
# sub-VP SDE scalars                                    lib/algorithms/advanced/sde_lib.py
# --------------------------------------------------------------------------------------

BETA_MIN = 0.1      # configs/default_pose_gen_configs.py:67-69
BETA_MAX = 20.0
NUM_SCALES = 1000   # sde.N, model.num_scales
T_START = 0.1       # config.model.t  (configs/optim/concat_pose_optimization_h36m.py:67)
SAMPLING_EPS = 0.01  # config.ZeDO.sampling_eps


def _exp32(a):
    """Correctly rounded float32 exp of a float32 argument.  1 - exp(.) cancels heavily at small t
    (std ~ 2e-3 at t = 0.01), so a 1-ulp difference between exp implementations (torch/SLEEF,
    numpy SIMD, glibc, CUDA) shows up as ~3e-5 relative in std: float32 noise of the reference."""
    return np.exp(np.asarray(a, dtype=np.float64)).astype(f32)


def subvp_sde_scalars(t, beta_min=BETA_MIN, beta_max=BETA_MAX):
    """beta(t), diffusion g(t) of ``subVPSDE.sde`` (sde_lib.py:187-192), float32 op order.

    Returns (beta_t, diffusion) as float32 (arrays if ``t`` is an array).
    """
    t = np.asarray(t, dtype=f32)
    b0, db = f32(beta_min), f32(beta_max - beta_min)
    beta_t = b0 + t * db
    discount = f32(1.0) - _exp32(f32(-2 * beta_min) * t - db * t ** 2)
    diffusion = np.sqrt(beta_t * discount, dtype=f32)
    return beta_t.astype(f32), diffusion.astype(f32)


def subvp_marginal_std(t, beta_min=BETA_MIN, beta_max=BETA_MAX):
    """``subVPSDE.marginal_prob`` std (sde_lib.py:194-198): 1 - exp(2*log_mean_coeff)."""
    t = np.asarray(t, dtype=f32)
    db = f32(beta_max - beta_min)
    lmc = f32(-0.25) * t ** 2 * db - f32(0.5) * t * f32(beta_min)
    return (f32(1.0) - _exp32(f32(2.0) * lmc)).astype(f32)


def vp_sde_scalars(t, beta_min=BETA_MIN, beta_max=BETA_MAX):
    """``VPSDE.sde`` (sde_lib.py:137-141): beta(t), sqrt(beta(t))."""
    t = np.asarray(t, dtype=f32)
    beta_t = f32(beta_min) + t * f32(beta_max - beta_min)
    return beta_t.astype(f32), np.sqrt(beta_t, dtype=f32)


def vp_marginal_std(t, beta_min=BETA_MIN, beta_max=BETA_MAX):
    """``VPSDE.marginal_prob`` std (sde_lib.py:143-147)."""
    t = np.asarray(t, dtype=f32)
    lmc = f32(-0.25) * t ** 2 * f32(beta_max - beta_min) - f32(0.5) * t * f32(beta_min)
    return np.sqrt(f32(1.0) - np.exp(f32(2.0) * lmc, dtype=f32), dtype=f32)


def oil_time_grid(steps=NUM_SCALES, t_start=T_START, eps=SAMPLING_EPS):
    """``torch.linspace(sde.T, sampling_eps, sample_num)`` (run/opt_main.py:197-198).

    torch's float32 linspace is symmetric: the first half is start + i*step, the second
    half is end - (steps-1-i)*step, with step = (end-start)/(steps-1) in float32 and the
    multiply-add fused (one rounding) -- bit-exact against torch 2.11 CPU.
    """
    start, end = f32(t_start), f32(eps)
    step = np.float64((end - start) / f32(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(start) + step * i).astype(f32)
    hi = (np.float64(end) - step * (steps - 1 - i)).astype(f32)
    return np.where(i < steps // 2, lo, hi).astype(f32)


# --------------------------------------------------------------------------------------
# score network                                          lib/algorithms/advanced/model.py
# --------------------------------------------------------------------------------------

def timestep_embedding(timesteps, embedding_dim=512, max_positions=10000):
    """``get_timestep_embedding`` (model.py:81-95): [sin(t*f), cos(t*f)], f_k = exp(-k ln(1e4)/(half-1))."""
    timesteps = np.atleast_1d(np.asarray(timesteps, dtype=f32))
    half = embedding_dim // 2
    coef = math.log(max_positions) / (half - 1)
    freqs = np.exp(np.arange(half, dtype=f32) * f32(-coef), dtype=f32)
    arg = timesteps[:, None] * freqs[None, :]
    emb = np.concatenate([np.sin(arg, dtype=f32), np.cos(arg, dtype=f32)], axis=1)
    if embedding_dim % 2 == 1:
        emb = np.pad(emb, ((0, 0), (0, 1)))
    return emb.astype(f32)


def silu(x):
    """``nn.SiLU``: x * sigmoid(x)."""
    return (x / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def group_norm(x, gamma, beta, groups=32, eps=1e-5):
    """``nn.GroupNorm(32, C)`` on a [B, C] tensor (model.py:116,145,150): groups of C/32
    contiguous channels, biased variance, affine."""
    B, C = x.shape
    xg = x.reshape(B, groups, C // groups)
    mean = xg.mean(axis=2, keepdims=True, dtype=f32)
    var = ((xg - mean) ** 2).mean(axis=2, keepdims=True, dtype=f32)
    y = (xg - mean) / np.sqrt(var + f32(eps), dtype=f32)
    return (y.reshape(B, C) * gamma[None, :] + beta[None, :]).astype(f32)


def linear(x, w, b):
    """``nn.Linear``: x @ w.T + b with w in the reference layout [out, in]."""
    return (x @ w.T + b[None, :]).astype(f32)


def log_f32(t) -> np.ndarray:
    """``torch.log`` of a float32 tensor (model.py:249).  torch's float32 logarithm is correctly rounded on the whole
    OIL time grid (checked against the imported reference by gen_golden.py) while numpy's float32 SIMD log is 1 ulp
    off on 46 of the 1000 grid points -- and the Fourier features turn 1 ulp of log t into 2e-4 -- so the logarithm
    is formed in float64 and rounded once."""
    return np.log(np.atleast_1d(np.asarray(t, dtype=f32)).astype(np.float64)).astype(f32)


def gaussian_fourier_projection(x, Wp) -> np.ndarray:
    """``GaussianFourierProjection.forward`` (model.py:27-36): x_proj = ((x * W) * 2) * pi in float32 (the python
    scalars are cast to the tensor dtype), [sin, cos]."""
    x = np.atleast_1d(np.asarray(x, dtype=f32))
    proj = ((x[:, None] * np.asarray(Wp, dtype=f32)[None, :]) * f32(2.0)) * f32(np.pi)
    return np.concatenate([np.sin(proj, dtype=f32), np.cos(proj, dtype=f32)], axis=1).astype(f32)


def time_embed(W: Weights, t999) -> np.ndarray:
    """temb = SiLU(shared_time_embed(emb)) (model.py:246-259); [n_t, embed_dim].  emb = posit_proj(t) for the
    'positional' embedding of every shipped optimisation config, gauss_proj(log t) for 'fourier' (the default of
    configs/default_pose_gen_configs.py:71), which is recognised by its state_dict entry ``gauss_proj.W``."""
    if "gauss_proj.W" in W:
        emb = gaussian_fourier_projection(log_f32(t999), W["gauss_proj.W"])
    else:
        emb = timestep_embedding(t999, W["shared_time_embed.0.weight"].shape[1])
    return silu(linear(emb, W["shared_time_embed.0.weight"], W["shared_time_embed.0.bias"]))


def score_forward(W: Weights, x: np.ndarray, t999, n_blocks=2) -> np.ndarray:
    """``ScoreModelFC_Adv.forward`` (model.py:215-298) in eval mode, scale_by_sigma=False.

    x: [B, J, 3] float32; t999: scalar or [B] time labels (= 999*t, utils.py:762).
    ``condition`` and ``mask`` are never read by the reference forward (every use is
    commented out, model.py:225-244) so they are not parameters here.
    """
    B = x.shape[0]
    h_in = x.reshape(B, -1).astype(f32)
    t999 = np.asarray(t999, dtype=f32)
    temb = time_embed(W, t999)  # [1 or B, E]
    if temb.shape[0] == 1:
        temb_b = temb  # broadcast row
    else:
        temb_b = temb

    def tproj(name):
        return linear(temb_b, W[name + ".weight"], W[name + ".bias"])

    h = linear(h_in, W["pre_dense.weight"], W["pre_dense.bias"]) + tproj("pre_dense_t")
    h = silu(group_norm(h, W["pre_gnorm.weight"], W["pre_gnorm.bias"]))
    for k in range(1, n_blocks + 1):
        h1 = linear(h, W[f"b{k}_dense1.weight"], W[f"b{k}_dense1.bias"]) + tproj(f"b{k}_dense1_t")
        h1 = silu(group_norm(h1, W[f"b{k}_gnorm1.weight"], W[f"b{k}_gnorm1.bias"]))
        h2 = linear(h1, W[f"b{k}_dense2.weight"], W[f"b{k}_dense2.bias"]) + tproj(f"b{k}_dense2_t")
        h2 = silu(group_norm(h2, W[f"b{k}_gnorm2.weight"], W[f"b{k}_gnorm2.bias"]))
        h = (h + h2).astype(f32)
    out = linear(h, W["post_dense.weight"], W["post_dense.bias"])
    return out.reshape(x.shape).astype(f32)


def control_score_forward(W: Weights, x: np.ndarray, t999, n_blocks=2) -> np.ndarray:
    """``Control_ScoreModelFC_Adv.forward`` (control_model.py:277-382), eval mode.

    Quirk kept on purpose (control_model.py:340-341): ``c = dense2_copy(c)`` is
    immediately overwritten by ``c = dense2_t_copy(temb)``, so ``dense2_copy`` and
    ``gnorm1_copy`` never influence the output and ``c2`` is batch-invariant.
    """
    B = x.shape[0]
    xb = x.reshape(B, -1).astype(f32)
    temb = time_embed(W, np.asarray(t999, dtype=f32))

    def lin(name, v):
        return linear(v, W[name + ".weight"], W[name + ".bias"])

    c = silu(lin("zc_layer_1", W["infant_cond"][None, :].astype(f32)))
    c = (xb + c).astype(f32)
    c = lin("pre_dense_copy", c) + lin("pre_dense_t_copy", temb)
    c0 = lin("zc_layer_2", c)
    c = silu(group_norm(c, W["pre_gnorm_copy.weight"], W["pre_gnorm_copy.bias"]))

    h = lin("pre_dense", xb) + lin("pre_dense_t", temb) + c0
    h = silu(group_norm(h, W["pre_gnorm.weight"], W["pre_gnorm.bias"]))
    for k in range(1, n_blocks + 1):
        orc = c
        c = lin(f"b{k}_dense1_copy", c) + lin(f"b{k}_dense1_t_copy", temb)
        c1 = lin(f"zc_b{k}_1", c)
        # gnorm1_copy / dense2_copy are dead: their result is overwritten below.
        c = lin(f"b{k}_dense2_t_copy", temb)
        c2 = lin(f"zc_b{k}_2", c)
        c = silu(group_norm(np.broadcast_to(c, (B, c.shape[1])).astype(f32),
                            W[f"b{k}_gnorm2_copy.weight"], W[f"b{k}_gnorm2_copy.bias"]))
        c = (orc + c).astype(f32)

        h1 = lin(f"b{k}_dense1", h) + lin(f"b{k}_dense1_t", temb) + c1
        h1 = silu(group_norm(h1, W[f"b{k}_gnorm1.weight"], W[f"b{k}_gnorm1.bias"]))
        h2 = lin(f"b{k}_dense2", h1) + lin(f"b{k}_dense2_t", temb) + c2
        h2 = silu(group_norm(h2, W[f"b{k}_gnorm2.weight"], W[f"b{k}_gnorm2.bias"]))
        h = (h + h2).astype(f32)
    return lin("post_dense", h).reshape(x.shape).astype(f32)


# --------------------------------------------------------------------------------------
# sampler                       lib/algorithms/advanced/sampling.py, sde_lib.py, utils.py
# --------------------------------------------------------------------------------------

def score_fn_subvp(W: Weights, x, t, forward=score_forward):
    """``get_score_fn`` sub-VP branch (utils.py:751-777): -model(x, 999 t)/std(t)."""
    t = f32(t)
    eps_theta = forward(W, x, t * f32(999))
    std = subvp_marginal_std(t)
    return (-eps_theta / std).astype(f32)


def reverse_sde_subvp(W: Weights, x, t, probability_flow=True, forward=score_forward):
    """``RSDE.sde`` (sde_lib.py:93-100). NB the score factor is 1.0 in *both* branches."""
    beta_t, g = subvp_sde_scalars(t)
    drift = f32(-0.5) * beta_t * x
    score = score_fn_subvp(W, x, t, forward)
    drift = drift - g ** 2 * score * f32(1.0)
    diffusion = f32(0.0) if probability_flow else g
    return drift.astype(f32), f32(diffusion)


def euler_maruyama_update(W: Weights, x, t, z=None, probability_flow=True, n_scales=NUM_SCALES,
                          forward=score_forward):
    """``EulerMaruyamaPredictor.update_fn`` (sampling.py:180-191).

    ``z`` is the injected ``randn_like(x)`` tensor; with probability_flow=True the
    diffusion is 0 and z is dead.  Returns (x, x_mean).
    """
    dt = f32(-1.0 / n_scales)
    drift, diffusion = reverse_sde_subvp(W, x, t, probability_flow, forward)
    x_mean = (x + drift * dt).astype(f32)
    if z is None:
        z = np.zeros_like(x)
    x_new = (x_mean + diffusion * f32(np.sqrt(-dt)) * z).astype(f32)
    return x_new, x_mean


def reverse_diffusion_update(W: Weights, x, t, z=None, probability_flow=True, n_scales=NUM_SCALES,
                             forward=score_forward):
    """``ReverseDiffusionPredictor.update_fn`` (sampling.py:195-205) with the default
    ``SDE.discretize`` (sde_lib.py:52-69) and ``RSDE.discretize`` (sde_lib.py:102-107)."""
    dt = f32(1.0 / n_scales)
    beta_t, g = subvp_sde_scalars(t)
    f = f32(-0.5) * beta_t * x * dt
    G = g * np.sqrt(dt, dtype=f32)
    rev_f = f - G ** 2 * score_fn_subvp(W, x, t, forward)
    rev_G = f32(0.0) if probability_flow else G
    x_mean = (x - rev_f).astype(f32)
    if z is None:
        z = np.zeros_like(x)
    return (x_mean + rev_G * z).astype(f32), x_mean


def langevin_update(W: Weights, x, t, noises: Sequence[np.ndarray], snr=0.16, n_steps=1,
                    forward=score_forward, norm_means: Optional[Tuple[float, float]] = None):
    """``LangevinCorrector.update_fn`` (sampling.py:258-287) for the sub-VP SDE.

    NB the reference reads ``sde.alphas`` which only ``VPSDE`` defines
    (sde_lib.py:125-127 vs 168-185): with the sub-VP SDE this corrector raises
    AttributeError, so alpha = 1 - linspace(b0/N, b1/N, N)[timestep] is the VPSDE
    definition applied here.  Step size uses the *batch mean* norms (sampling.py:281-283).
    """
    N = NUM_SCALES
    timestep = int(np.floor(float(f32(t) * f32(N - 1) / f32(T_START))))
    betas = np.linspace(BETA_MIN / N, BETA_MAX / N, N).astype(f32)
    alpha = f32(1.0) - betas[min(timestep, N - 1)]
    x_mean = x
    for i in range(n_steps):
        grad = score_fn_subvp(W, x, t, forward)
        noise = noises[i].astype(f32)
        if norm_means is None:
            gn = np.linalg.norm(grad.reshape(grad.shape[0], -1), axis=-1).astype(f32).mean(dtype=f32)
            nn_ = np.linalg.norm(noise.reshape(noise.shape[0], -1), axis=-1).astype(f32).mean(dtype=f32)
        else:
            gn, nn_ = f32(norm_means[0]), f32(norm_means[1])
        step = (f32(snr) * nn_ / gn) ** 2 * f32(2) * alpha
        x_mean = (x + step * grad).astype(f32)
        x = (x_mean + np.sqrt(step * f32(2), dtype=f32) * noise).astype(f32)
    return x, x_mean


def score_fn_vp(W: Weights, x, t, forward=score_forward):
    """``get_score_fn`` VP branch, continuous (utils.py:751-777): -model(x, 999 t) / sqrt(1 - exp(2 lmc))."""
    t = f32(t)
    return (-forward(W, x, t * f32(999)) / vp_marginal_std(t)).astype(f32)


def ve_sigma(t, sigma_min=0.01, sigma_max=50.0):
    """``VESDE.marginal_prob`` std (sde_lib.py:241-244): sigma_min (sigma_max / sigma_min)^t."""
    return (f32(sigma_min) * np.power(f32(sigma_max / sigma_min), f32(t), dtype=f32)).astype(f32)


def score_fn_ve(W: Weights, x, t, forward=score_forward):
    """``get_score_fn`` VE branch, continuous (utils.py:779-795): model(x, labels = sigma(t)), no rescale."""
    return forward(W, x, ve_sigma(t)).astype(f32)


def _timestep(t, n_scales=NUM_SCALES, T=T_START):
    """``(t * (sde.N - 1) / sde.T).long()`` in float32 (sampling.py:234,273)."""
    return int(np.trunc(f32(f32(t) * f32(n_scales - 1)) / f32(T)))


def vp_discrete_betas(beta_min=BETA_MIN, beta_max=BETA_MAX, n_scales=NUM_SCALES):
    """``VPSDE.discrete_betas`` (sde_lib.py:125): torch.linspace(beta_min / N, beta_max / N, N), float32."""
    return oil_time_grid(n_scales, beta_min / n_scales, beta_max / n_scales)


def ancestral_update_vp(W: Weights, x, t, z, forward=score_forward, T=T_START):
    """``AncestralSamplingPredictor.vpsde_update_fn`` (sampling.py:233-241)."""
    beta = vp_discrete_betas()[_timestep(t, T=T)]
    score = score_fn_vp(W, x, t, forward)
    x_mean = ((x + beta * score) / np.sqrt(f32(1.0) - beta, dtype=f32)).astype(f32)
    return (x_mean + np.sqrt(beta, dtype=f32) * z).astype(f32), x_mean


def ancestral_update_ve(W: Weights, x, t, z, forward=score_forward, T=T_START, sigma_min=0.01, sigma_max=50.0):
    """``AncestralSamplingPredictor.vesde_update_fn`` (sampling.py:220-231)."""
    sig = np.exp(np.linspace(np.log(sigma_min), np.log(sigma_max), NUM_SCALES).astype(f32), dtype=f32)
    ts = _timestep(t, T=T)
    sigma, adj = sig[ts], (f32(0.0) if ts == 0 else sig[ts - 1])
    score = score_fn_ve(W, x, t, forward)
    x_mean = (x + score * (sigma ** 2 - adj ** 2)).astype(f32)
    std = np.sqrt((adj ** 2 * (sigma ** 2 - adj ** 2)) / (sigma ** 2), dtype=f32)
    return (x_mean + std * z).astype(f32), x_mean


def langevin_update_vp(W: Weights, x, t, noises: Sequence[np.ndarray], snr=0.16, n_steps=1, forward=score_forward,
                       T=T_START, norm_means: Optional[Tuple[float, float]] = None, ald=False):
    """``LangevinCorrector.update_fn`` (sampling.py:258-287) / ``AnnealedLangevinDynamics.update_fn`` (:290-324,
    ``ald=True``) for the VP SDE (the only SDE of the three that defines ``alphas``, sde_lib.py:126)."""
    alpha = f32(1.0) - vp_discrete_betas()[_timestep(t, T=T)]
    std_m = vp_marginal_std(t)
    x_mean = x
    for i in range(n_steps):
        grad = score_fn_vp(W, x, t, forward)
        noise = noises[i].astype(f32)
        if ald:
            step = (f32(snr) * std_m) ** 2 * f32(2) * alpha
        else:
            if norm_means is None:
                gn = np.linalg.norm(grad.reshape(grad.shape[0], -1), axis=-1).astype(f32).mean(dtype=f32)
                nn_ = np.linalg.norm(noise.reshape(noise.shape[0], -1), axis=-1).astype(f32).mean(dtype=f32)
            else:
                gn, nn_ = f32(norm_means[0]), f32(norm_means[1])
            step = (f32(snr) * nn_ / gn) ** 2 * f32(2) * alpha
        x_mean = (x + step * grad).astype(f32)
        x = (x_mean + np.sqrt(step * f32(2), dtype=f32) * noise).astype(f32)
    return x, x_mean


def pc_sampler_step(W: Weights, denoise_x, t, denoise=True, probability_flow=True, z=None,
                    forward=score_forward):
    """One call of ``pc_sampler`` (sampling.py:450-527) for the shipped configuration
    predictor='euler_maruyama', corrector='none' (NoneCorrector returns (x, x)).

    Returns (trajs [1,B,J,3] = x_mean, results [B,J,3] = x_mean if denoise else x).
    ``condition``, ``gradient``, ``t_step`` and ``args`` are ignored by the reference
    (sampling.py:491-499: mask*0, x = denoise_x, ``t_step < 0`` never true).
    """
    x, x_mean = euler_maruyama_update(W, denoise_x.astype(f32), t, z, probability_flow, forward=forward)
    trajs = x_mean[None].copy()
    return trajs, (x_mean if denoise else x)


# --------------------------------------------------------------------------------------
# geometry                              lib/algorithms/advanced/simple_zeroshot_opt.py
# --------------------------------------------------------------------------------------

def inv3x3(M):
    """Batched 3x3 inverse (the reference uses ``torch.inverse``, simple_zeroshot_opt.py:61,92)."""
    return np.linalg.inv(M.astype(f32)).astype(f32)


def gradient_field(key2d, key3d, K, t=None, conf=None):
    """``gradient_field_gen`` (simple_zeroshot_opt.py:46-125), noise_type=None.

    key2d [B,J,2], key3d [B,J,3], K [B,3,3], conf [B,J] or None (clamped IN PLACE to
    [1e-4, 1] like the reference, :64-66), t [B,1,3] or None.
    Returns (gradient [B,J,3], T [B,1,3]).  When ``t`` is None the translation is the
    least-squares solve of :73-93 (rows weighted conf^2 on both A and b, sign flip if
    T_z < 0); otherwise T = t.
    """
    key2d = key2d.astype(f32)
    key3d = key3d.astype(f32)
    B, J, _ = key3d.shape
    Kinv = inv3x3(K)
    h2d = np.concatenate([key2d, np.ones((B, J, 1), f32)], axis=-1)
    if conf is not None:
        conf[conf > 1] = 1
        conf[conf < 1e-4] = 1e-4
    ray = np.einsum("bij,bnj->bni", Kinv, h2d).astype(f32)
    ray = (ray / ray[:, :, 2:]).astype(f32)
    if t is None:
        A = np.zeros((B, 2 * J, 3), f32)
        b = np.zeros((B, 2 * J, 1), f32)
        b[:, 0::2, :] = key3d[:, :, 0:1] - key3d[:, :, 2:3] * ray[:, :, 0:1]
        b[:, 1::2, :] = key3d[:, :, 1:2] - key3d[:, :, 2:3] * ray[:, :, 1:2]
        A[:, 0::2, 0] = -1
        A[:, 0::2, 2] = ray[:, :, 0]
        A[:, 1::2, 1] = -1
        A[:, 1::2, 2] = ray[:, :, 1]
        if conf is not None:
            w = (conf[:, :, None] * conf[:, :, None]).astype(f32)
            A[:, 0::2, :] *= w
            A[:, 1::2, :] *= w
            b[:, 0::2, :] *= w
            b[:, 1::2, :] *= w
        At = np.transpose(A, (0, 2, 1))
        ATA = (At @ A).astype(f32)
        ATb = (At @ b).astype(f32)
        T = np.transpose(inv3x3(ATA) @ ATb, (0, 2, 1)).astype(f32)
        neg = T[:, :, 2] < 0
        T[neg] = T[neg] * -1
    else:
        T = t.astype(f32)
    ray = (ray / np.linalg.norm(ray, axis=-1, keepdims=True)).astype(f32)
    point = key3d + T
    proj = np.sum(point * ray, axis=-1, keepdims=True, dtype=f32) * ray  # perpendicular_distance :33-36
    return (proj - point).astype(f32), T


# --------------------------------------------------------------------------------------
# IPO: rotation / scale fit           simple_zeroshot_opt.py:8-31, run/opt_main.py:175-195
# --------------------------------------------------------------------------------------

AXES = "xyz"


def quaternion_to_matrix(q):
    """``quaternion_to_matrix`` (utils.py:59-88); q [...,4] real part first, NOT normalised
    by the caller: two_s = 2 / sum(q^2)."""
    q = q.astype(f32)
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = f32(2.0) / (q * q).sum(-1, dtype=f32)
    o = np.stack([
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
    ], axis=-1)
    return o.reshape(q.shape[:-1] + (3, 3)).astype(f32)


def init_translation(key2d, K, ipo_T, pelvis=(0, 0)):
    """T0 = IPO_T * normalize(K^-1 [u_pelvis, v_pelvis, 1]) (run/opt_main.py:177-179); [B,1,3].
    pelvis = (a, b): the pelvis pixel is (key2d[a] + key2d[b]) / 2 -- joint 0, or joints 0 and 3 for
    SyRIP (run/opt_main_infant.py:259-262)."""
    B = key2d.shape[0]
    pix = ((key2d[:, pelvis[0], :2].astype(f32) + key2d[:, pelvis[1], :2].astype(f32)) / f32(2)).astype(f32)
    pelvis = np.concatenate([pix, np.ones((B, 1), f32)], axis=-1)
    T = np.einsum("bij,bj->bi", inv3x3(K), pelvis)[:, None, :].astype(f32)
    return (T / np.linalg.norm(T, axis=-1, keepdims=True) * f32(ipo_T)).astype(f32)


def ray_init(key2d, K, T, pelvis=(0, 0)):
    """Infant driver's initial pose (run/opt_main_infant.py:281-292): back-projected 2D rays scaled so
    the pelvis ray has length |T|, pelvis-subtracted.  key2d [B,J,2], T [B,1,3] -> [B,J,3]."""
    B, J = key2d.shape[:2]
    h2d = np.concatenate([key2d.astype(f32), np.ones((B, J, 1), f32)], axis=-1)
    ray = np.einsum("bij,bnj->bni", inv3x3(K), h2d).astype(f32)
    root = ((ray[:, pelvis[0]:pelvis[0] + 1] + ray[:, pelvis[1]:pelvis[1] + 1]) / f32(2)).astype(f32)
    ray = (ray / np.linalg.norm(root, axis=-1, keepdims=True)).astype(f32)
    ray = (ray * np.linalg.norm(T, axis=-1, keepdims=True)).astype(f32)
    root = ((ray[:, pelvis[0]:pelvis[0] + 1] + ray[:, pelvis[1]:pelvis[1] + 1]) / f32(2)).astype(f32)
    return (ray - root).astype(f32)


def init_hypothesis(cluster_poses, sid, B):
    """x0 = ones_like(gt) * (sample_poses - sample_poses[:,0:1])[sid] (run/opt_main.py:167-173)."""
    rel = (cluster_poses - cluster_poses[:, 0:1, :]).astype(f32)
    return np.broadcast_to(rel[sid:sid + 1], (B,) + rel.shape[1:]).astype(f32).copy()


def ipo_project(q, scale, x_key, T0, K, minT, maxT):
    """``RotOpt.forward`` (simple_zeroshot_opt.py:20-25): uv = proj(K (R x + T0*clamp(scale)))."""
    R = quaternion_to_matrix(q)
    s = np.clip(scale, f32(minT), f32(maxT)).astype(f32)
    p = np.einsum("bij,bkj->bki", R, x_key) + T0 * s[:, None, None]
    P = np.einsum("bij,bkj->bki", K.astype(f32), p.astype(f32)).astype(f32)
    return (P[:, :, :2] / P[:, :, 2:]).astype(f32), p.astype(f32), P


def ipo_loss_and_grad(q, scale, x_key, uv_key, T0, K, minT, maxT, axes_mask, b_global=None):
    """Analytic gradient of ``torch.mean(L1Loss(reduction='none')(rot2d, cond))``
    (run/opt_main.py:186-191) w.r.t. (q, scale); SURVEY appendix B.3.

    axes_mask: bool[4] over (w, x, y, z); w is always trainable (rot_vect), the others
    only if listed in config.ZeDO.RotAxes.  b_global: batch size used in the mean
    (global batch when the poses are sharded).
    """
    B, nk, _ = x_key.shape
    bg = B if b_global is None else b_global
    lam = f32(1.0 / (bg * nk * 2))
    uv, p, P = ipo_project(q, scale, x_key, T0, K, minT, maxT)
    diff = uv - uv_key.astype(f32)
    loss = np.abs(diff).sum(dtype=np.float64) * float(lam)
    d_uv = (np.sign(diff) * lam).astype(f32)
    du, dv = d_uv[..., 0], d_uv[..., 1]
    Pz = P[..., 2]
    dP = np.stack([du / Pz, dv / Pz, -(du * P[..., 0] + dv * P[..., 1]) / (Pz * Pz)], axis=-1).astype(f32)
    dp = np.einsum("bji,bkj->bki", K.astype(f32), dP).astype(f32)  # K^T dP
    G = np.einsum("bki,bkj->bij", dp, x_key.astype(f32)).astype(f32)  # sum_k dp_k x_k^T
    inside = ((scale >= f32(minT)) & (scale <= f32(maxT))).astype(f32)
    d_scale = (np.einsum("bki,bi->b", dp, T0[:, 0, :]) * inside).astype(f32)

    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    n = (q * q).sum(-1, dtype=f32)
    s2 = f32(2.0) / n
    zero = np.zeros_like(w)
    A = np.stack([
        -(y * y + z * z), x * y - z * w, x * z + y * w,
        x * y + z * w, -(x * x + z * z), y * z - x * w,
        x * z - y * w, y * z + x * w, -(x * x + y * y)], axis=-1).reshape(B, 3, 3)
    dA = [
        np.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1),
        np.stack([zero, y, z, y, -2 * x, -w, z, w, -2 * x], -1),
        np.stack([-2 * y, x, w, x, zero, z, -w, z, -2 * y], -1),
        np.stack([-2 * z, -w, x, w, -2 * z, y, x, y, zero], -1),
    ]
    GA = (G * A).sum(axis=(1, 2), dtype=f32)
    d_q = np.zeros_like(q)
    for a in range(4):
        if not axes_mask[a]:
            continue
        GdA = (G.reshape(B, 9) * dA[a]).sum(-1, dtype=f32)
        d_q[:, a] = (f32(-4.0) * q[:, a] / (n * n)) * GA + s2 * GdA
    return loss, d_q.astype(f32), d_scale


def axes_to_mask(rot_axes: str):
    return np.array([True] + [a in rot_axes for a in AXES])


def adam_update(theta, g, m, v, step, lr=0.1, b1=0.9, b2=0.999, eps=1e-8):
    """``torch.optim.Adam`` single-tensor update (defaults, no weight decay, no amsgrad)."""
    m[...] = f32(b1) * m + f32(1 - b1) * g
    v[...] = f32(b2) * v + f32(1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = f32(lr / bc1)
    denom = (np.sqrt(v, dtype=f32) / f32(math.sqrt(bc2))) + f32(eps)
    theta[...] = theta - step_size * (m / denom)


def ipo_fit(x0, key2d, K, keylist, rot_axes, ipo_T, minT, maxT, iters=500, b_global=None,
            trace=None):
    """The IPO loop of run/opt_main.py:175-195 with the analytic gradient.

    Returns (R [B,3,3], T [B,1,3]) where T = T0*clamp(scale).  ``trace`` (optional list)
    receives (q, scale) copies after every iteration.
    """
    B = x0.shape[0]
    T0 = init_translation(key2d, K, ipo_T)
    mask = axes_to_mask(rot_axes)
    q = np.zeros((B, 4), f32)
    q[:, 0] = 1
    scale = np.ones((B,), f32)
    mq, vq = np.zeros_like(q), np.zeros_like(q)
    ms, vs = np.zeros_like(scale), np.zeros_like(scale)
    xk = x0[:, keylist, :].astype(f32)
    uvk = key2d[:, keylist, :2].astype(f32)
    for it in range(1, iters + 1):
        _, dq, ds = ipo_loss_and_grad(q, scale, xk, uvk, T0, K, minT, maxT, mask, b_global)
        adam_update(q, dq, mq, vq, it)
        q[:, ~mask] = 0  # frozen components never move (they are not parameters)
        adam_update(scale, ds, ms, vs, it)
        if trace is not None:
            trace.append((q.copy(), scale.copy()))
    R = quaternion_to_matrix(q)
    T = (T0 * np.clip(scale, f32(minT), f32(maxT))[:, None, None]).astype(f32)
    return R, T


# --------------------------------------------------------------------------------------
# OIL loop                                                      run/opt_main.py:197-222
# --------------------------------------------------------------------------------------

def oil_loop_schedule(W: Weights, x, T, key2d, K, conf, ts, phase_switch, dump_steps=(), forward=score_forward):
    """steps x {gradient_field_gen -> x += g -> pc_sampler} (run/opt_main.py:202-220) over an explicit
    time schedule ``ts``.  Steps i < phase_switch keep the IPO translation; afterwards T is re-solved
    each step and carried.  Returns (x_final, T_final, {step: pose after that step})."""
    x = x.astype(f32).copy()
    T = T.astype(f32).copy()
    dumps = {}
    want = set(int(d) for d in dump_steps)
    for i in range(len(ts)):
        if i < phase_switch:
            g, _ = gradient_field(key2d, x, K, t=T, conf=conf)
        else:
            g, T = gradient_field(key2d, x, K, t=None, conf=conf)
        x = (x + g).astype(f32)
        _, x = pc_sampler_step(W, x, ts[i], forward=forward)
        if i in want:
            dumps[i] = x.copy()
    return x, T, dumps


def oil_loop(W: Weights, x, T, key2d, K, conf, steps=NUM_SCALES, t_start=T_START, eps=SAMPLING_EPS,
             phase_div=5, dump_every=0, forward=score_forward):
    """The shipped configuration of the loop: ts = linspace(T, eps, steps), phase switch at steps // 5
    (run/opt_main.py:197-206).  Returns (x_final, T_final, [(i, pose)] every dump_every steps)."""
    ts = oil_time_grid(steps, t_start, eps)
    want = [i for i in range(steps) if dump_every and (i % dump_every == dump_every - 1 or i == 0)]
    x, T, d = oil_loop_schedule(W, x, T, key2d, K, conf, ts, steps // phase_div, want, forward)
    return x, T, sorted(d.items())


def run_hypothesis(W: Weights, cluster_poses, sid, key2d_conf, K, cfg, fixed_RT=None,
                   steps=None, forward=score_forward):
    """One pass of the ``for sid in range(args.hypo)`` body (run/opt_main.py:166-222).

    key2d_conf: db_2d [B,J,3] = (u, v, conf).  cfg: dict with the ZeDO block fields.
    fixed_RT: optional (R, T) to bypass the (chaotic) IPO fit.
    """
    B = key2d_conf.shape[0]
    key2d = key2d_conf[:, :, :2].astype(f32)
    conf = key2d_conf[:, :, 2].astype(f32).copy()
    x0 = init_hypothesis(cluster_poses, sid, B)
    if fixed_RT is None:
        R, T = ipo_fit(x0, key2d, K, cfg["IPO_keylist"], cfg["RotAxes"], cfg["IPO_T"],
                       cfg["IPO_minScaleT"], cfg["IPO_maxScaleT"], cfg["IPO_iterations"])
    else:
        R, T = fixed_RT
    x = np.einsum("bij,bnj->bni", R, x0).astype(f32)
    n = cfg["OIL_iterations"] if steps is None else steps
    x, T, _ = oil_loop(W, x, T, key2d, K, conf, steps=n, forward=forward)
    return x, T, R


# --------------------------------------------------------------------------------------
# evaluation                         lib/utils/transforms.py:42-148, lib/dataset/h36m.py:365-442
# --------------------------------------------------------------------------------------

def procrustes_align(pose, pose_gt):
    """``align_to_gt`` = ``procrustes(pose_gt, pose)[1]`` (transforms.py:42-127,143-148):
    scaling=True, reflection='best' (no determinant fix -> reflections allowed)."""
    A = np.array(pose_gt, copy=True)
    Bm = np.array(pose, copy=True)
    A_bar, B_bar = A.mean(0), Bm.mean(0)
    A0, B0 = A - A_bar, Bm - B_bar
    A_norm, B_norm = np.sqrt((A0 ** 2).sum()), np.sqrt((B0 ** 2).sum())
    A0 = A0 / A_norm
    B0 = B0 / B_norm
    U, s, Vt = np.linalg.svd(A0.T @ B0)
    R = Vt.T @ U.T
    return A_norm * s.sum() * (B0 @ R) + A_bar


def mpjpe(pred, gt):
    """mean_j ||pred_j - gt_j||_2 (h36m.py:406-407)."""
    return np.mean(np.sqrt(np.square(pred - gt).sum(axis=1)))


def compute_pck(gts, preds, eval_joints=None, threshold=150):
    """``compute_PCK`` (utils.py:814-834): % of joints whose error (mm, scale 1000) is < threshold."""
    if eval_joints is None:
        eval_joints = list(range(gts.shape[1]))
    err = np.sqrt(np.sum(np.power(preds - gts, 2), axis=2))[:, eval_joints] * 1000
    return float((err < threshold).sum() / err.size) * 100


def compute_auc(gts, preds, eval_joints=None):
    """``compute_AUC`` (utils.py:837-848): mean PCK over thresholds linspace(0, 150, 31)."""
    return float(np.mean([compute_pck(gts, preds, eval_joints, t) for t in np.linspace(0, 150, 31)]))


def hypothesis_std(preds):
    """Diversity report of mpii3dHP.py:487-490: ``multi_preds_cam - multi_preds_cam[:, :, [0], :]``, root
    dropped, ``[..., c].std(axis=1).mean()`` for c = x, y, z.  preds [N,S,J,3] -> 3 floats."""
    rel = (preds - preds[:, :, [0], :])[:, :, 1:, :]
    return tuple(float(rel[..., c].std(axis=1).mean()) for c in range(3))


def eval_multi(preds, gts, protocol2=False, actions=None, joint_subset=None, valid_ind=None):
    """Multi-hypothesis evaluation (h36m.py:365-442; pw3d.py:286-345 for the plain mean).

    preds [N,S,J,3]; gts [N,J,3] already root-relative in metres.  Per pose: error of
    every hypothesis (optionally after Procrustes), ``np.argmin`` / ``np.amin`` over S.
    actions: optional int[N] in 2..16 -> H36M aggregate = mean over the 15 action means
    (h36m.py:424-433); otherwise the plain mean over poses.
    valid_ind: optional per-pose collections of admissible hypothesis indices (h36m.py:399-401): the
    others are skipped, and the returned index counts inside the kept list, as ``np.argmin`` of the
    reference's filtered ``multi_results`` does.
    Returns (aggregate, per_pose_min [N], argmin [N]).
    """
    N, S = preds.shape[:2]
    res = np.zeros(N, dtype=np.result_type(preds.dtype, gts.dtype))
    idx = np.zeros(N, dtype=np.int64)
    for n in range(N):
        gt = gts[n]
        errs = []
        for s in range(S):
            if valid_ind is not None and s not in valid_ind[n]:
                continue
            pred = preds[n, s]
            if protocol2:
                pred = procrustes_align(pred, gt)
            if joint_subset is not None:
                errs.append(mpjpe(pred[joint_subset], gt[joint_subset]))
            else:
                errs.append(mpjpe(pred, gt))
        idx[n] = int(np.argmin(errs))
        res[n] = np.amin(errs)
    if actions is not None:
        agg = float(np.mean([np.mean(res[actions == a]) for a in range(2, 17)]))
    else:
        agg = float(np.mean(res))
    return agg, res, idx


# --------------------------------------------------------------------------------------
# synthetic inputs (shared by tests, bench and the golden generator)   SURVEY.md 8(d)
# --------------------------------------------------------------------------------------

H36M_SKELETON = [[0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [0, 7], [7, 8], [8, 9], [9, 10],
                 [8, 11], [11, 12], [12, 13], [8, 14], [14, 15], [15, 16]]  # h36m.py:445-448

_H36M_TEMPLATE = np.array([
    [0.00, 0.00, 0.00], [-0.13, 0.00, 0.00], [-0.13, 0.44, 0.00], [-0.13, 0.88, 0.00],
    [0.13, 0.00, 0.00], [0.13, 0.44, 0.00], [0.13, 0.88, 0.00], [0.00, -0.24, 0.00],
    [0.00, -0.48, 0.00], [0.00, -0.58, 0.00], [0.00, -0.70, 0.00], [0.17, -0.44, 0.00],
    [0.30, -0.20, 0.00], [0.32, 0.04, 0.00], [-0.17, -0.44, 0.00], [-0.30, -0.20, 0.00],
    [-0.32, 0.04, 0.00]], dtype=f32)  # metres, y down (camera frame), root = pelvis


def kmeans_lloyd(x, init_centres, iters):
    """Lloyd's k-means as the cluster-file generator runs it (zedo_release_b200/clusters.py; the reference ships the
    cluster files, not their generator -- run/opt_main.py:58-65 loads them, run/opt_main_infant.py:25,34 only imports
    scipy.cluster.vq / sklearn KMeans).  x [N, D] float32, float64 distances and means, ties to the lowest index,
    an empty cluster keeps its centre.  Returns (centres [S, D] float32, labels [N], squared distances [N] float64)."""
    x64 = x.astype(np.float64)
    c = init_centres.astype(f32).copy()

    def assign(c):
        d = ((x64[:, None, :] - c.astype(np.float64)[None]) ** 2).sum(-1)
        lab = d.argmin(1)
        return lab, d[np.arange(len(x)), lab]

    for _ in range(iters):
        lab, _ = assign(c)
        for k in range(len(c)):
            m = lab == k
            if m.any():
                c[k] = x64[m].mean(0).astype(f32)
    lab, dist = assign(c)
    return c, lab.astype(np.int32), dist


def skeleton_template(n_joints=17):
    if n_joints == 17:
        return _H36M_TEMPLATE.copy()
    rng = np.random.default_rng(99)
    t = rng.normal(0, 0.3, (n_joints, 3)).astype(f32)
    t[0] = 0
    return t


def make_weights(seed=0, n_joints=17, hidden=1024, embed=512, n_blocks=2, control=False, fourier=False) -> Weights:
    """Random-init weights with the reference's state_dict names and the default
    ``nn.Linear`` init range U(-1/sqrt(fan_in), 1/sqrt(fan_in)); GroupNorm affine is
    randomised (U(0.5,1.5), U(-0.2,0.2)) so the affine path is exercised.  Generated with
    numpy's PCG64 so the same seed gives the same weights on every machine."""
    rng = np.random.default_rng(seed)
    D = n_joints * 3
    W: Weights = {}

    def lin(name, fin, fout):
        bound = 1.0 / math.sqrt(fin)
        W[name + ".weight"] = rng.uniform(-bound, bound, (fout, fin)).astype(f32)
        W[name + ".bias"] = rng.uniform(-bound, bound, (fout,)).astype(f32)

    def gn(name):
        W[name + ".weight"] = rng.uniform(0.5, 1.5, (hidden,)).astype(f32)
        W[name + ".bias"] = rng.uniform(-0.2, 0.2, (hidden,)).astype(f32)

    lin("pre_dense", D, hidden)
    lin("pre_dense_t", embed, hidden)
    gn("pre_gnorm")
    lin("shared_time_embed.0", embed, embed)
    for k in range(1, n_blocks + 1):
        lin(f"b{k}_dense1", hidden, hidden)
        lin(f"b{k}_dense1_t", embed, hidden)
        gn(f"b{k}_gnorm1")
        lin(f"b{k}_dense2", hidden, hidden)
        lin(f"b{k}_dense2_t", embed, hidden)
        gn(f"b{k}_gnorm2")
    lin("post_dense", hidden, D)
    if control:
        W["infant_cond"] = rng.normal(0, 1, (D,)).astype(f32)
        lin("zc_layer_1", D, D)
        lin("zc_layer_2", hidden, hidden)
        lin("pre_dense_copy", D, hidden)
        lin("pre_dense_t_copy", embed, hidden)
        gn("pre_gnorm_copy")
        for k in range(1, n_blocks + 1):
            lin(f"zc_b{k}_1", hidden, hidden)
            lin(f"zc_b{k}_2", hidden, hidden)
            lin(f"b{k}_dense1_copy", hidden, hidden)
            lin(f"b{k}_dense1_t_copy", embed, hidden)
            gn(f"b{k}_gnorm1_copy")
            lin(f"b{k}_dense2_copy", hidden, hidden)
            lin(f"b{k}_dense2_t_copy", embed, hidden)
            gn(f"b{k}_gnorm2_copy")
    if fourier:  # GaussianFourierProjection(embed_dim, scale=30): W ~ N(0, 30^2), drawn last so the other tensors
        W["gauss_proj.W"] = (rng.normal(0, 1, (embed // 2,)) * 30.0).astype(f32)  # match the positional variant
    return W


def make_synthetic_dataset(n_poses, n_joints=17, seed=1234, detected_2d=True, n_clusters=1,
                           dtype_gt=np.float32):
    """Synthetic H36M-format inputs: returns a dict with
    db_3d [N,J,3] (root-relative metres), db_2d [N,J,3]=(u,v,conf), camera_param [N,3,3],
    actions [N] in 2..16, clusters [S,J,3], root [N,3]."""
    rng = np.random.default_rng(seed)
    tmpl = skeleton_template(n_joints)
    ang = rng.uniform(-np.pi, np.pi, n_poses)
    c, s = np.cos(ang), np.sin(ang)
    Ry = np.zeros((n_poses, 3, 3))
    Ry[:, 0, 0], Ry[:, 0, 2], Ry[:, 1, 1], Ry[:, 2, 0], Ry[:, 2, 2] = c, s, 1, -s, c
    gt = np.einsum("bij,nj->bni", Ry, tmpl) + rng.normal(0, 0.05, (n_poses, n_joints, 3))
    gt = gt - gt[:, 0:1]
    root = np.stack([rng.uniform(-1, 1, n_poses), rng.uniform(-1, 1, n_poses),
                     rng.uniform(3, 7, n_poses)], axis=-1)
    cam = gt + root[:, None, :]
    K = np.zeros((n_poses, 3, 3))
    K[:, 0, 0] = 1145.0 + rng.uniform(-5, 5, n_poses)
    K[:, 1, 1] = 1145.0 + rng.uniform(-5, 5, n_poses)
    K[:, 0, 2] = 512.0 + rng.uniform(-4, 4, n_poses)
    K[:, 1, 2] = 515.0 + rng.uniform(-4, 4, n_poses)
    K[:, 2, 2] = 1.0
    proj = np.einsum("bij,bnj->bni", K, cam)
    uv = proj[:, :, :2] / proj[:, :, 2:]
    if detected_2d:
        uv = uv + rng.normal(0, 5.0, uv.shape)
        conf = rng.uniform(0.3, 1.0, (n_poses, n_joints))
    else:
        conf = np.ones((n_poses, n_joints))
    clusters = tmpl[None] + rng.normal(0, 0.1, (n_clusters, n_joints, 3))
    clusters = clusters - clusters[:, 0:1]
    return dict(
        db_3d=gt.astype(dtype_gt),
        db_2d=np.concatenate([uv, conf[..., None]], axis=-1).astype(f32),
        camera_param=K.astype(f32),
        actions=(2 + np.arange(n_poses) % 15).astype(np.int64),
        clusters=clusters.astype(f32),
        root=root.astype(f32),
    )


PW3D_ORDER = [5, 2, 6, 3, 11, 14, 12, 15, 13, 16, 1, 4, 8, 10, 0, 7, 9]  # lib/dataset/pw3d.py:76


def h36m_items_from_arrays(ds, dtype=np.float64):
    """The synthetic dataset as the list of ground-truth items ``h36m_<subset>.pkl`` holds
    (lib/dataset/h36m.py:205-234): ``joint_3d_camera`` [17,3] millimetres (absolute), ``joint_3d_image``
    [17,3] (u, v, depth), ``camera_param`` {fx, fy, cx, cy} as 0-d arrays, ``image_path``, ``action``."""
    items = []
    cam_mm = (ds["db_3d"].astype(np.float64) + ds["root"].astype(np.float64)[:, None, :]) * 1000.0
    for n in range(len(cam_mm)):
        K = ds["camera_param"][n]
        img = np.concatenate([ds["db_2d"][n, :, :2], cam_mm[n, :, 2:3]], axis=-1)
        items.append({"joint_3d_camera": cam_mm[n].astype(dtype), "joint_3d_image": img.astype(dtype),
                      "camera_param": {"fx": np.array(K[0, 0]), "fy": np.array(K[1, 1]), "cx": np.array(K[0, 2]),
                                       "cy": np.array(K[1, 2])},
                      "image_path": f"synthetic/{n:08d}.jpg", "action": int(ds["actions"][n])})
    return items


def pw3d_npz_from_arrays(ds):
    """The synthetic dataset under the keys of ``pw3d_<subset>.npz`` (lib/dataset/pw3d.py:184-199): joints in the
    file's own order (``order_change`` maps file joint i to H36M joint PW3D_ORDER[i]), root-relative + ``root_cam``."""
    N = len(ds["db_3d"])
    rel = np.empty((N, 17, 3), np.float64)
    for i in range(17):
        rel[:, i] = ds["db_3d"][:, PW3D_ORDER[i]]
    K = ds["camera_param"].astype(np.float64)
    return {"keypoints3d17_relative": rel, "root_cam": ds["root"].astype(np.float64),
            "cam_param": np.array({"f": np.stack([K[:, 0, 0], K[:, 1, 1]], -1), "c": np.stack([K[:, 0, 2], K[:, 1, 2]], -1)},
                                  dtype=object),
            "image_width": np.full(N, 1920), "image_height": np.full(N, 1080),
            "image_path": np.array([f"synthetic/{n:08d}.jpg" for n in range(N)])}


H36M_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=[0, 1, 4], RotAxes="z", IPO_T=3,
                     IPO_minScaleT=0.5, IPO_maxScaleT=2, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_h36m.py:70-81
PW3D_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="z", IPO_T=8,
                     IPO_minScaleT=0.2, IPO_maxScaleT=2, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_pw3d.py:72-81
SKI_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="y", IPO_T=20,
                    IPO_minScaleT=0.5, IPO_maxScaleT=2, OIL_iterations=1000,
                    sampling_eps=0.01)  # configs/optim/concat_pose_optimization_ski.py:72-81
MINI_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(17)), RotAxes="xyz", IPO_T=1,
                     IPO_minScaleT=0, IPO_maxScaleT=4, OIL_iterations=1000,
                     sampling_eps=0.01)  # configs/optim/concat_pose_optimization_mini.py:73-85
SYRIP_ZEDO_CFG = dict(IPO_iterations=500, IPO_keylist=list(range(12)), RotAxes="xyz", IPO_T=1,
                      IPO_minScaleT=0.5, IPO_maxScaleT=8, OIL_iterations=1000,
                      sampling_eps=0.01)  # configs/optim/concat_pose_optimization_syrip.py:73-86
