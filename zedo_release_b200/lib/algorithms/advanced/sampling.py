"""Sampler factory with the reference's interface (lib/algorithms/advanced/sampling.py).

``get_sampling_fn(config, sde, shape, inverse_scaler, eps, device)`` returns ``pc_sampler`` whose
call signature and return value -- ``(trajs np.float32 [1,B,J,3], results np.float32 [B,J,3])`` --
are the reference's (sampling.py:450-527).  For the configuration every shipped config uses
(sub-VP SDE, Euler-Maruyama or reverse-diffusion predictor, 'none' corrector, our
``ScoreModelFC_Adv``) one call is a single ``zedo_sde_step``: bias-table build, six fused layer
kernels and the fused predictor update.  Other registered predictors/correctors compose the
generic ``update_fn`` objects below around the same CUDA score network.
"""
import abc
import functools

import numpy as np
import torch

from zedo_release_b200.parallel import global_batch_mean, global_sum_
from . import sde_lib
from . import utils as mutils
from .utils import from_flattened_numpy, to_flattened_numpy, get_score_fn  # noqa: F401

_CORRECTORS = {}
_PREDICTORS = {}


def _make_register(table, what):
    def register(cls=None, *, name=None):
        def _register(c):
            key = c.__name__ if name is None else name
            if key in table:
                raise ValueError(f'Already registered model with name: {key}')
            table[key] = c
            return c
        return _register if cls is None else _register(cls)
    register.__doc__ = f"A decorator for registering {what} classes."
    return register


register_predictor = _make_register(_PREDICTORS, "predictor")
register_corrector = _make_register(_CORRECTORS, "corrector")


def get_predictor(name):
    return _PREDICTORS[name]


def get_corrector(name):
    return _CORRECTORS[name]


def get_sampling_fn(config, sde, shape, inverse_scaler, eps, device=None):
    """'pc' -> predictor-corrector sampler, 'ode' -> black-box probability-flow ODE; anything else
    raises ValueError (sampling.py:80-127)."""
    if device is None:
        device = config.device
    name = config.sampling.method.lower()
    if name == 'ode':
        return get_ode_sampler(sde=sde, shape=shape, inverse_scaler=inverse_scaler,
                               denoise=config.sampling.noise_removal, eps=eps, device=device)
    if name == 'pc':
        return get_pc_sampler(sde=sde, shape=shape,
                              predictor=get_predictor(config.sampling.predictor.lower()),
                              corrector=get_corrector(config.sampling.corrector.lower()),
                              inverse_scaler=inverse_scaler, snr=config.sampling.snr,
                              n_steps=config.sampling.n_steps_each,
                              probability_flow=config.sampling.probability_flow,
                              continuous=config.training.continuous, denoise=config.sampling.noise_removal,
                              eps=eps, device=device)
    raise ValueError(f"Sampler name {config.sampling.method} unknown.")


class Predictor(abc.ABC):
    """A predictor: ``update_fn(x, t, condition, mask) -> (x, x_mean)`` (sampling.py:130-153)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn

    @abc.abstractmethod
    def update_fn(self, x, t, condition, mask):
        pass


class Corrector(abc.ABC):
    """A corrector: ``update_fn(x, t, condition, mask) -> (x, x_mean)`` (sampling.py:156-177)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.sde, self.score_fn, self.snr, self.n_steps = sde, score_fn, snr, n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, condition, mask):
        pass


@register_predictor(name='euler_maruyama')
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, x, t, condition, mask):
        dt = -1. / self.rsde.N
        z = torch.randn_like(x)
        drift, diffusion = self.rsde.sde(x, t, condition, mask)
        x_mean = x + drift * dt
        return x_mean + diffusion[:, None, None] * np.sqrt(-dt) * z, x_mean


@register_predictor(name='reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
    def update_fn(self, x, t, condition, mask):
        f, G = self.rsde.discretize(x, t, condition, mask)
        z = torch.randn_like(x)
        x_mean = x - f
        return x_mean + G[:, None, None] * z, x_mean


def _fused_score(score_fn, x, t):
    """The pieces of a fused update, or None when the eager composition has to serve the call: the packed plan of
    the network behind ``score_fn``, the time label and the std divisor of get_score_fn (utils.py:751-795) for a
    batch-uniform time.  Float32 scalars are formed on the host exactly as the reference forms them on tensors."""
    info = getattr(score_fn, "zedo", None)
    if info is None or info["train"] or not hasattr(info["model"], "zedo_plan") or not x.is_cuda:
        return None
    sde, model, continuous = info["sde"], info["model"], info["continuous"]
    if getattr(model.config.model, "scale_by_sigma", False) or not bool((t == t[0]).all()):
        return None
    t0 = t[:1].detach().cpu().float()
    if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
        if not (continuous or isinstance(sde, sde_lib.subVPSDE)):
            return None  # discrete VP labels: not a shipped configuration
        label = float((t0 * 999)[0])
        std_div = float(sde.marginal_prob(torch.zeros(1, 1, 1), t0)[1][0])
    elif isinstance(sde, sde_lib.VESDE) and continuous:
        label = float(sde.marginal_prob(torch.zeros(1, 1, 1), t0)[1][0])
        std_div = 0.0
    else:
        return None
    return model.zedo_plan(x.shape[0]), label, std_div, t0, getattr(model, "gemm_mode", None)


@register_predictor(name='ancestral_sampling')
class AncestralSamplingPredictor(Predictor):
    """Ancestral sampling; VE/VP SDEs only, no probability flow (sampling.py:208-244)."""

    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__(sde, score_fn, probability_flow)
        if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
        assert not probability_flow, "Probability flow not supported by ancestral sampling"

    def update_fn(self, x, t, condition, mask):
        sde = self.sde
        fused = _fused_score(self.score_fn, x, t)
        if fused is not None:  # network forward + one fused update kernel, noise injected by the caller's RNG
            plan, label, std_div, t0, mode = fused
            ts = (t0 * (sde.N - 1) / sde.T).long()
            noise = torch.randn_like(x)
            plan.score_stats(x, label, std_div=std_div, mode=mode)
            if isinstance(sde, sde_lib.VESDE):
                sigma = sde.discrete_sigmas[ts]
                adjacent = torch.where(ts == 0, torch.zeros_like(t0), sde.discrete_sigmas[ts - 1])
                return plan.noise_update("ancestral_ve", x, noise, std_div, float(sigma[0]), float(adjacent[0]))
            return plan.noise_update("ancestral_vp", x, noise, std_div, float(sde.discrete_betas[ts][0]))
        timestep = (t * (sde.N - 1) / sde.T).long()
        score = self.score_fn(x, t, condition, mask)
        noise = torch.randn_like(x)
        if isinstance(sde, sde_lib.VESDE):
            sigma = sde.discrete_sigmas[timestep]
            adjacent = torch.where(timestep == 0, torch.zeros_like(t), sde.discrete_sigmas.to(t.device)[timestep - 1])
            x_mean = x + score * (sigma ** 2 - adjacent ** 2)[:, None, None]
            std = torch.sqrt((adjacent ** 2 * (sigma ** 2 - adjacent ** 2)) / (sigma ** 2))
            return x_mean + std[:, None, None] * noise, x_mean
        beta = sde.discrete_betas.to(t.device)[timestep]
        x_mean = (x + beta[:, None, None] * score) / torch.sqrt(1. - beta)[:, None, None]
        return x_mean + torch.sqrt(beta)[:, None, None] * noise, x_mean


@register_predictor(name='none')
class NonePredictor(Predictor):
    """An empty predictor that does nothing."""

    def __init__(self, sde, score_fn, probability_flow=False):
        pass

    def update_fn(self, x, t, condition, mask):
        return x, x


def _check_langevin_sde(sde):
    if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
        raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")


def _langevin_alpha(sde, t):
    if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
        timestep = (t * (sde.N - 1) / sde.T).long()
        return sde.alphas.to(t.device)[timestep]  # like the reference: AttributeError for subVPSDE
    return torch.ones_like(t)


@register_corrector(name='langevin')
class LangevinCorrector(Corrector):
    """Langevin corrector; the step size uses the BATCH MEAN of the gradient / noise norms
    (sampling.py:281-283)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        _check_langevin_sde(sde)

    def update_fn(self, x, t, condition, mask):
        alpha = _langevin_alpha(self.sde, t)
        x_mean = x
        fused = _fused_score(self.score_fn, x, t)
        if fused is not None:
            plan, label, std_div, t0, mode = fused
            a0 = float(_langevin_alpha(self.sde, t0)[0])
            for _ in range(self.n_steps):
                noise = torch.randn_like(x)
                stats = plan.score_stats(x, label, z=noise, std_div=std_div, want_stats=True, mode=mode)
                global_sum_(stats)  # batch means over ALL ranks when the poses are sharded (one 3-double all_reduce)
                x, x_mean = plan.noise_update("langevin", x, noise, std_div, self.snr, a0, stats=stats)
            return x, x_mean
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, condition, mask)
            noise = torch.randn_like(x)
            # batch means: taken over ALL ranks when the poses are sharded (one 2-float all_reduce each)
            grad_norm = global_batch_mean(torch.norm(grad.reshape(grad.shape[0], -1), dim=-1))
            noise_norm = global_batch_mean(torch.norm(noise.reshape(noise.shape[0], -1), dim=-1))
            step_size = (self.snr * noise_norm / grad_norm) ** 2 * 2 * alpha
            x_mean = x + step_size[:, None, None] * grad
            x = x_mean + torch.sqrt(step_size * 2)[:, None, None] * noise
        return x, x_mean


@register_corrector(name='ald')
class AnnealedLangevinDynamics(Corrector):
    """Annealed Langevin dynamics of NCSN/NCSNv2 (sampling.py:290-324)."""

    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        _check_langevin_sde(sde)

    def update_fn(self, x, t, condition, mask):
        alpha = _langevin_alpha(self.sde, t)
        std = self.sde.marginal_prob(x, t)[1]
        x_mean = x
        fused = _fused_score(self.score_fn, x, t)
        if fused is not None:
            plan, label, std_div, t0, mode = fused
            a0 = float(_langevin_alpha(self.sde, t0)[0])
            std_m = float(self.sde.marginal_prob(torch.zeros(1, 1, 1), t0)[1][0])
            for _ in range(self.n_steps):
                noise = torch.randn_like(x)
                plan.score_stats(x, label, std_div=std_div, mode=mode)
                x, x_mean = plan.noise_update("ald", x, noise, std_div, self.snr, a0, std_m)
            return x, x_mean
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, condition, mask)
            noise = torch.randn_like(x)
            step_size = (self.snr * std) ** 2 * 2 * alpha
            x_mean = x + step_size[:, None, None] * grad
            x = x_mean + noise * torch.sqrt(step_size * 2)[:, None, None]
        return x, x_mean


@register_corrector(name='none')
class NoneCorrector(Corrector):
    """An empty corrector that does nothing."""

    def __init__(self, sde, score_fn, snr, n_steps):
        pass

    def update_fn(self, x, t, condition, mask):
        return x, x


def shared_predictor_update_fn(x, t, condition, mask, sde, model, predictor, probability_flow, continuous):
    score_fn = mutils.get_score_fn(sde, model, train=False, continuous=continuous)
    obj = NonePredictor(sde, score_fn, probability_flow) if predictor is None else predictor(sde, score_fn,
                                                                                             probability_flow)
    return obj.update_fn(x, t, condition, mask)


def shared_corrector_update_fn(x, t, condition, mask, sde, model, corrector, continuous, snr, n_steps):
    score_fn = mutils.get_score_fn(sde, model, train=False, continuous=continuous)
    obj = NoneCorrector(sde, score_fn, snr, n_steps) if corrector is None else corrector(sde, score_fn, snr, n_steps)
    return obj.update_fn(x, t, condition, mask)


_FUSED_PREDICTORS = {EulerMaruyamaPredictor: "euler_maruyama", ReverseDiffusionPredictor: "reverse_diffusion"}


def _fused_step_available(sde, model, predictor, corrector):
    return (type(sde) is sde_lib.subVPSDE and predictor in _FUSED_PREDICTORS
            and corrector in (None, NoneCorrector) and hasattr(model, "zedo_plan")
            and not getattr(model.config.model, "scale_by_sigma", False))


def get_pc_sampler(sde, shape, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False,
                   continuous=False, denoise=True, eps=1e-3, device='cuda'):
    """Create the single-step PC sampler of the reference (sampling.py:400-529)."""
    predictor_update_fn = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor,
                                            probability_flow=probability_flow, continuous=continuous)
    corrector_update_fn = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector,
                                            continuous=continuous, snr=snr, n_steps=n_steps)

    def pc_sampler(model, condition, gradient=None, denoise_x=None, t=None, t_step=None, args=None):
        """ONE corrector + predictor update starting from ``denoise_x``.  ``condition``, ``gradient``,
        ``args`` are accepted and ignored exactly like the reference (mask * 0, x = denoise_x);
        ``t_step < 0`` forces t = 1.  Returns (trajs [1,B,J,3], results [B,J,3]) as numpy."""
        with torch.no_grad():
            x = denoise_x
            batch_size = condition.shape[0]
            t_val = torch.as_tensor(t)
            if t_step is not None and t_step < 0:
                t_val = torch.ones_like(t_val)
            if _fused_step_available(sde, model, predictor, corrector):
                plan = model.zedo_plan(x.shape[0])
                z = None if probability_flow else torch.randn_like(x)
                x, x_mean = plan.sde_step(x, float(t_val), z=z, predictor=_FUSED_PREDICTORS[predictor],
                                          probability_flow=probability_flow, beta_min=sde.beta_0,
                                          beta_max=sde.beta_1, n_scales=sde.N,
                                          mode=getattr(model, "gemm_mode", None))
            else:
                mask = torch.ones_like(x) * 0
                vec_t = torch.ones(batch_size, device=t_val.device) * t_val
                x, x_mean = corrector_update_fn(x, vec_t, condition, mask, model=model)
                x, x_mean = predictor_update_fn(x, vec_t, condition, mask, model=model)
            x_mean = x_mean.cpu().numpy()
            trajs = x_mean[None].copy()  # trajs[-1] = x_mean (sampling.py:524-526)
            return trajs, (x_mean if denoise else x.cpu().numpy())

    return pc_sampler


def get_ode_sampler(sde, shape, inverse_scaler, denoise=False, rtol=1e-5, atol=1e-5, method='RK45', eps=1e-3,
                    device='cuda'):
    """Probability-flow ODE sampler over scipy's black-box solver (sampling.py:532-603).  The
    reference's version cannot run (its drift_fn is called without ``condition``, :587 vs :561);
    this one passes ``None`` for condition/mask, which the score network ignores anyway."""
    from scipy import integrate

    def drift_fn(model, x, t):
        score_fn = get_score_fn(sde, model, train=False, continuous=True)
        return sde.reverse(score_fn, probability_flow=True).sde(x, t, None, None)[0]

    def ode_sampler(model, z=None):
        with torch.no_grad():
            x = sde.prior_sampling(shape).to(device) if z is None else z

            def ode_func(t, flat):
                xx = from_flattened_numpy(flat, shape).to(device).type(torch.float32)
                vec_t = torch.ones(shape[0], device=xx.device) * t
                return to_flattened_numpy(drift_fn(model, xx, vec_t))

            sol = integrate.solve_ivp(ode_func, (sde.T, eps), to_flattened_numpy(x), rtol=rtol, atol=atol,
                                      method=method)
            x = torch.tensor(sol.y[:, -1]).reshape(shape).to(device).type(torch.float32)
            if denoise:
                score_fn = get_score_fn(sde, model, train=False, continuous=True)
                vec_eps = torch.ones(x.shape[0], device=x.device) * eps
                _, x = ReverseDiffusionPredictor(sde, score_fn, probability_flow=False).update_fn(x, vec_eps, None, None)
            return inverse_scaler(x), sol.nfev

    return ode_sampler
