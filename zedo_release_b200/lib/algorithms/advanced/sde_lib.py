"""SDE objects with the reference's interface (lib/algorithms/advanced/sde_lib.py).

They carry the SDE *parameters* and the scalar noise schedules.  The per-pose arithmetic of the
shipped configuration (sub-VP, probability flow) runs inside ``zedo_sde_step`` / ``zedo_oil_loop``
(csrc/geom.cu), which re-derives the same float32 scalars on the host (csrc/api.cu: subvp_coef).
The tensor methods exist for interface parity -- ``.sde``, ``.marginal_prob``, ``.discretize``,
``.reverse`` -- and work on [B]-sized schedule vectors.

Layout of this module (differs from the reference, the public names do not):
``_ReverseTime`` is one module-level class instead of a class built inside ``SDE.reverse``;
``_LinearBeta`` holds the beta(t) = beta_0 + t (beta_1 - beta_0) schedule shared by VP and sub-VP.
"""
import abc
import math

import torch


def _bcast(v):
    """[B] schedule vector -> [B,1,1] for broadcasting against poses [B,J,3]."""
    return v[:, None, None]


def _standard_normal_logp(z, variance=1.0):
    dims = math.prod(z.shape[1:])
    return -0.5 * dims * math.log(2 * math.pi * variance) - z.pow(2).sum(dim=(1, 2, 3)) / (2 * variance)


class SDE(abc.ABC):
    """Forward SDE dx = f(x,t) dt + g(t) dw on t in [0, T], discretised in N steps."""

    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        """End time."""

    @abc.abstractmethod
    def sde(self, x, t):
        """(drift f(x,t) [B,J,3], diffusion g(t) [B])."""

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        """(mean, std) of p_t(x | x_0 = x)."""

    @abc.abstractmethod
    def prior_sampling(self, shape):
        """A sample of p_T."""

    @abc.abstractmethod
    def prior_logp(self, z):
        """log p_T(z)."""

    def discretize(self, x, t):
        """One Euler-Maruyama step of the forward SDE: (f dt, g sqrt(dt)) with dt = 1/N
        (reference sde_lib.py:52-69)."""
        step = 1 / self.N
        f, g = self.sde(x, t)
        return f * step, g * torch.sqrt(torch.tensor(step, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE, or the probability-flow ODE (reference sde_lib.py:71-109)."""
        return _ReverseTime(self, score_fn, probability_flow)


class _ReverseTime:
    """dx = [f - g^2 score] dt (+ g dw) run backwards in time.  NB the reference multiplies the score
    by 1.0 in BOTH modes (sde_lib.py:97,105) -- not by 0.5 for the ODE as upstream score_sde does."""

    def __init__(self, forward, score_fn, probability_flow):
        self._fwd = forward
        self._score = score_fn
        self.N = forward.N
        self.probability_flow = probability_flow

    @property
    def T(self):
        return self._fwd.T

    def __getattr__(self, name):  # beta_0, sigma_max, discrete_betas, ... of the wrapped SDE
        return getattr(self.__dict__["_fwd"], name)

    def sde(self, x, t, condition, mask):
        f, g = self._fwd.sde(x, t)
        f = f - _bcast(g) ** 2 * self._score(x, t, condition, mask)
        return f, (torch.zeros(1, device=f.device) if self.probability_flow else g)

    def discretize(self, x, t, condition, mask):
        f, G = self._fwd.discretize(x, t)
        f = f - _bcast(G) ** 2 * self._score(x, t, condition, mask)
        return f, (torch.zeros_like(G) if self.probability_flow else G)


class _LinearBeta(SDE):
    """beta(t) = beta_0 + t (beta_1 - beta_0) and log of the mean coefficient of the VP family."""

    def __init__(self, beta_min, beta_max, N, T):
        super().__init__(N)
        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self._T = T

    @property
    def T(self):
        return self._T

    def _beta(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def _log_mean_coeff(self, t):
        return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        return _standard_normal_logp(z)


class VPSDE(_LinearBeta):
    """Variance-preserving SDE (reference sde_lib.py:112-165), with the DDPM tables."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(beta_min, beta_max, N, T)
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = self.alphas_cumprod.sqrt()
        self.sqrt_1m_alphas_cumprod = (1. - self.alphas_cumprod).sqrt()

    def sde(self, x, t):
        b = self._beta(t)
        return -0.5 * _bcast(b) * x, b.sqrt()

    def marginal_prob(self, x, t):
        c = self._log_mean_coeff(t)
        return torch.exp(_bcast(c)) * x, torch.sqrt(1. - torch.exp(2. * c))

    def discretize(self, x, t):
        """DDPM ancestral discretisation: f = (sqrt(alpha_i) - 1) x, G = sqrt(beta_i)."""
        i = (t * (self.N - 1) / self.T).long()
        a = self.alphas.to(x.device)[i]
        return _bcast(a.sqrt()) * x - x, self.discrete_betas.to(x.device)[i].sqrt()


class subVPSDE(_LinearBeta):
    """The SDE of every shipped config (training.sde = 'subvpsde', reference sde_lib.py:168-206):
    g(t)^2 = beta(t) (1 - exp(-2 beta_0 t - (beta_1 - beta_0) t^2)),  std(t) = 1 - exp(2 log_mean_coeff)."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(beta_min, beta_max, N, T)

    def sde(self, x, t):
        b = self._beta(t)
        damp = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * _bcast(b) * x, torch.sqrt(b * damp)

    def marginal_prob(self, x, t):
        c = self._log_mean_coeff(t)
        return _bcast(torch.exp(c)) * x, 1 - torch.exp(2. * c)


class VESDE(SDE):
    """Variance-exploding SDE (reference sde_lib.py:209-261): sigma(t) = sigma_min (sigma_max/sigma_min)^t."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, T=1):
        super().__init__(N)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self._T = T
        self.discrete_sigmas = torch.exp(torch.linspace(math.log(sigma_min), math.log(sigma_max), N))

    @property
    def T(self):
        return self._T

    def _sigma(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def sde(self, x, t):
        rate = torch.sqrt(torch.tensor(2 * (math.log(self.sigma_max) - math.log(self.sigma_min)), device=t.device))
        return torch.zeros_like(x), self._sigma(t) * rate

    def marginal_prob(self, x, t):
        return x, self._sigma(t)

    def prior_sampling(self, shape):
        return self.sigma_max * torch.randn(*shape)

    def prior_logp(self, z):
        return _standard_normal_logp(z, self.sigma_max ** 2)

    def discretize(self, x, t):
        """SMLD / NCSN discretisation: G_i = sqrt(sigma_i^2 - sigma_{i-1}^2), sigma_{-1} = 0."""
        i = (t * (self.N - 1) / self.T).long()
        cur = self.discrete_sigmas.to(t.device)[i]
        prev = torch.where(i == 0, torch.zeros_like(t), self.discrete_sigmas[i - 1].to(t.device))
        return torch.zeros_like(x), torch.sqrt(cur ** 2 - prev ** 2)
