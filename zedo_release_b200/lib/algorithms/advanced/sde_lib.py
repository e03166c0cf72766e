"""SDE classes with the reference's interface (lib/algorithms/advanced/sde_lib.py).

These objects carry the SDE *parameters* and the scalar schedules; the per-pose arithmetic of the
shipped configuration (sub-VP, probability flow) runs inside ``zedo_sde_step`` / ``zedo_oil_loop``
(csrc/geom.cu: sde_update_kernel), which re-derives the same float32 scalars on the host.  The
tensor methods below exist for interface parity (``.sde``, ``.marginal_prob``, ``.discretize``,
``.reverse``) and operate on [B]-sized schedule vectors.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """Abstract SDE; ``N`` = number of discretisation steps (sde_lib.py:7-69)."""

    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        """End time of the SDE."""

    @abc.abstractmethod
    def sde(self, x, t):
        """Drift and diffusion of the forward SDE."""

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        """Mean and std of p_t(x)."""

    @abc.abstractmethod
    def prior_sampling(self, shape):
        """One sample from p_T."""

    @abc.abstractmethod
    def prior_logp(self, z):
        """log p_T(z)."""

    def discretize(self, x, t):
        """x_{i+1} = x_i + f_i(x_i) + G_i z_i, Euler-Maruyama by default (sde_lib.py:52-69)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE / probability-flow ODE (sde_lib.py:71-109).  NB the score factor is 1.0
        in both branches (not 0.5 for the ODE as in upstream score_sde)."""
        N, T = self.N, self.T
        fwd_sde, fwd_discretize = self.sde, self.discretize

        class RSDE(self.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, t, condition, mask):
                drift, diffusion = fwd_sde(x, t)
                score = score_fn(x, t, condition, mask)
                drift = drift - diffusion[:, None, None] ** 2 * score
                if self.probability_flow:
                    diffusion = torch.zeros(1, device=drift.device)
                return drift, diffusion

            def discretize(self, x, t, condition, mask):
                f, G = fwd_discretize(x, t)
                rev_f = f - G[:, None, None] ** 2 * score_fn(x, t, condition, mask)
                rev_G = torch.zeros_like(G) if self.probability_flow else G
                return rev_f, rev_G

        return RSDE()


def _gauss_logp(z, var=1.0):
    n = np.prod(z.shape[1:])
    return -n / 2. * np.log(2 * np.pi * var) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * var)


class VPSDE(SDE):
    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(N)
        self.beta_0, self.beta_1, self._T = beta_min, beta_max, T
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    @property
    def T(self):
        return self._T

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(lmc[:, None, None]) * x, torch.sqrt(1. - torch.exp(2. * lmc))

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        return _gauss_logp(z)

    def discretize(self, x, t):
        """DDPM discretisation (sde_lib.py:156-165)."""
        timestep = (t * (self.N - 1) / self.T).long()
        beta = self.discrete_betas.to(x.device)[timestep]
        alpha = self.alphas.to(x.device)[timestep]
        return torch.sqrt(alpha)[:, None, None] * x - x, torch.sqrt(beta)


class subVPSDE(SDE):
    """The SDE every shipped config uses (training.sde = 'subvpsde'; sde_lib.py:168-206)."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000, T=1):
        super().__init__(N)
        self.beta_0, self.beta_1, self._T = beta_min, beta_max, T

    @property
    def T(self):
        return self._T

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t * discount)

    def marginal_prob(self, x, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(lmc)[:, None, None] * x, 1 - torch.exp(2. * lmc)

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        return _gauss_logp(z)


class VESDE(SDE):
    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, T=1):
        super().__init__(N)
        self.sigma_min, self.sigma_max, self._T = sigma_min, sigma_max, T
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    @property
    def T(self):
        return self._T

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        scale = torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)), device=t.device))
        return torch.zeros_like(x), sigma * scale

    def marginal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max

    def prior_logp(self, z):
        return _gauss_logp(z, self.sigma_max ** 2)

    def discretize(self, x, t):
        """SMLD (NCSN) discretisation (sde_lib.py:251-260)."""
        timestep = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas.to(t.device)[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), self.discrete_sigmas[timestep - 1].to(t.device))
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)
