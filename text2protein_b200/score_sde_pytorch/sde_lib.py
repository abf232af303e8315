# API surface and closed-form expressions follow score_sde_pytorch as vendored by the reference
# (szhan227/text2protein, score_sde_pytorch/sde_lib.py):
# Copyright 2020 The Google Research Authors.
#
# Licensed under the Apache License, Version 2.0 (the "License");
# you may not use this file except in compliance with the License.
# You may obtain a copy of the License at
#
#     http://www.apache.org/licenses/LICENSE-2.0
#
# Unless required by applicable law or agreed to in writing, software
# distributed under the License is distributed on an "AS IS" BASIS,
# WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
# See the License for the specific language governing permissions and
# limitations under the License.
"""Forward SDEs and their reverse-time discretisations -- host-side mirror of the reference
``score_sde_pytorch/sde_lib.py`` (same class names, constructor arguments, attributes and method signatures).

Only the small per-step scalars are computed here (with the same torch expressions as the reference, so the
sigma / beta tables are bit-identical: sde_lib.py:118-122,210); the per-element update runs in the fused CUDA
step kernels (csrc/pc_step.cu) called from ``sampling.py``.
"""
import abc

import numpy as np
import torch


class SDE(abc.ABC):
    """Abstract SDE on mini-batches (reference sde_lib.py:7-103)."""

    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        """End time of the SDE."""

    @abc.abstractmethod
    def sde(self, x, t, context=None):
        """Returns (drift, diffusion) of the forward SDE."""

    @abc.abstractmethod
    def marginal_prob(self, x, t):
        """Returns (mean, std) of p_t(x | x_0)."""

    @abc.abstractmethod
    def prior_sampling(self, shape):
        """One sample from p_T."""

    @abc.abstractmethod
    def prior_logp(self, z):
        """log p_T(z)."""

    def discretize(self, x, t, context=None):
        """Euler-Maruyama: x_{i+1} = x_i + f_i(x_i) + G_i z_i (reference :49-64)."""
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t, context)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow=False):
        """Reverse-time SDE / probability-flow ODE (reference :66-103)."""
        fwd = self

        class RSDE(self.__class__):
            def __init__(self):  # deliberately skips the parent constructor, like the reference
                self.N = fwd.N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return fwd.T

            def sde(self, x, t, context=None):
                drift, diffusion = fwd.sde(x, t, context)
                score = score_fn(x, t, context)
                half = 0.5 if self.probability_flow else 1.
                drift = drift - diffusion[:, None, None, None] ** 2 * score * half
                return drift, (0. if self.probability_flow else diffusion)

            def discretize(self, x, t, context=None):
                f, G = fwd.discretize(x, t, context)
                half = 0.5 if self.probability_flow else 1.
                rev_f = f - G[:, None, None, None] ** 2 * score_fn(x, t, context) * half
                return rev_f, (torch.zeros_like(G) if self.probability_flow else G)

        return RSDE()


class VPSDE(SDE):
    """Variance-preserving SDE (reference :106-157)."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

    @property
    def T(self):
        return 1

    def sde(self, x, t, context=None):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None, None, None] * x, torch.sqrt(beta_t)

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(log_mean_coeff[:, None, None, None]) * x, torch.sqrt(1. - torch.exp(2. * log_mean_coeff))

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

    def timestep(self, t):
        return (t * (self.N - 1) / self.T).long()

    def discretize(self, x, t, context=None):
        """DDPM discretisation (reference :149-157)."""
        timestep = self.timestep(t)
        beta = self.discrete_betas.to(x.device)[timestep]
        alpha = self.alphas.to(x.device)[timestep]
        return torch.sqrt(alpha)[:, None, None, None] * x - x, torch.sqrt(beta)


class subVPSDE(SDE):
    """Present for API compatibility only: the reference class cannot run through the sampler (its ``sde`` lacks
    the ``context`` argument and it has no ``alphas``), so using it here raises as well."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N

    @property
    def T(self):
        return 1

    def sde(self, x, t, context=None):
        raise NotImplementedError("subVPSDE is not supported on the PC sampling path")

    def marginal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(log_mean_coeff)[:, None, None, None] * x, 1 - torch.exp(2. * log_mean_coeff)

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.


class VESDE(SDE):
    """Variance-exploding SDE (reference :199-245)."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(N)
        self.sigma_min, self.sigma_max, self.N = sigma_min, sigma_max, N
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    @property
    def T(self):
        return 1

    def sde(self, x, t, context=None):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        scale = torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)), device=t.device))
        return torch.zeros_like(x), sigma * scale

    def marginal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_sampling(self, shape):
        return torch.randn(*shape) * self.sigma_max

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2. * np.log(2 * np.pi * self.sigma_max ** 2) \
            - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)

    def timestep(self, t):
        return (t * (self.N - 1) / self.T).long()

    def discretize_G(self, t):
        """G_i = sqrt(sigma_i^2 - sigma_{i-1}^2) of the SMLD discretisation, without materialising f = 0.
        Tables stay on the host and are indexed with a host index (the reference re-uploads the table and
        indexes a CPU tensor with a device tensor every step, sde_lib.py:240-242)."""
        timestep = self.timestep(t.detach().cpu())
        sigma = self.discrete_sigmas[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(sigma), self.discrete_sigmas[timestep - 1])
        return torch.sqrt(sigma ** 2 - adjacent ** 2).to(t.device)

    def discretize(self, x, t, context=None):
        """SMLD (NCSN) discretisation (reference :237-245)."""
        return torch.zeros_like(x), self.discretize_G(t)
