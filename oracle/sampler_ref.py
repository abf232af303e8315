"""Oracle (test infrastructure): torch-CPU restatement of the reverse-SDE predictor-corrector sampler.

Follows score_sde_pytorch/sampling.py:157-289 and sde_lib.py:106-157,199-245 including the dtype promotions
of the reference (the score arrives as float64, the update runs in float64 and is cast back with
``.float()`` after each mask application).  Noise is injected through ``noise_fn(stream, like)`` so that a
run can be replayed with the exact normals the CUDA kernel generated.  Pinned against the imported reference
by tests/test_oracle_golden.py.
"""
import numpy as np
import torch


class VESDERef:
    # sde_lib.py:199-245
    def __init__(self, sigma_min=0.01, sigma_max=50.0, N=1000):
        self.sigma_min, self.sigma_max, self.N, self.T = sigma_min, sigma_max, N, 1
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))  # fp32 asc.

    def discretize_G(self, t):
        timestep = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), self.discrete_sigmas[timestep - 1])
        return torch.sqrt(sigma ** 2 - adjacent ** 2)

    def labels(self, t):
        # models/utils.py:166-169
        lab = self.T - t
        lab = lab * (self.N - 1)
        return torch.round(lab).long()


class VPSDERef:
    # sde_lib.py:106-157
    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000):
        self.beta_0, self.beta_1, self.N, self.T = beta_min, beta_max, N, 1
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)


def build_condition(x, condition):
    """sampling.py:260-275: returns (x, conditional_mask[bool B,C,N,N]); dict order matters."""
    cm = torch.ones_like(x).bool()
    if condition is not None:
        for k, v in condition.items():
            if k == "length":
                x = x * v.unsqueeze(1)
                cm = cm * v.unsqueeze(1)
                x[:, -1] = v
                cm[:, -1] = False
            elif k == "ss":
                x[:, 4:7] = v
                cm[:, 4:7] = False
            elif k == "inpainting":
                cm = cm * v["mask_inpaint"].unsqueeze(1)
                x = torch.where(cm, x, v["coords_6d"])
    return x, cm


@torch.no_grad()
def pc_sampler_ref(sde, score_model, shape, snr, n_steps=1, probability_flow=False, denoise=True, eps=1e-5,
                   condition=None, context=None, noise_fn=None, num_iters=None, x0=None):
    """``score_model(x, labels, context) -> float64 [B,C,N,N]`` (e.g. oracle.unet_ref.unet_forward bound to a
    state_dict).  ``noise_fn(stream, like) -> tensor`` supplies every normal draw; ``num_iters`` truncates the
    loop to the first K of sde.N iterations (bench / parity at small K).  VESDE only (every shipped config)."""
    from .philox_ref import STREAM_PRIOR, stream_corrector, stream_predictor

    assert isinstance(sde, VESDERef)
    B = shape[0]
    if x0 is None:
        x = noise_fn(STREAM_PRIOR, torch.empty(shape)) * sde.sigma_max  # sde_lib.py:229-230
    else:
        x = x0.clone()
    timesteps = torch.linspace(sde.T, eps, sde.N)
    x, cm = build_condition(x, condition)
    x_initial = x.detach().clone()
    x_mean = x
    K = sde.N if num_iters is None else num_iters
    for i in range(K):
        vec_t = torch.ones(B) * timesteps[i]
        labels = sde.labels(vec_t.clone())
        # ---- Langevin corrector, sampling.py:179-199 (alpha = 1 for VE)
        alpha = torch.ones_like(vec_t)
        for j in range(n_steps):
            grad = score_model(x, labels, context)
            noise = noise_fn(stream_corrector(i, j, n_steps), x)
            grad_norm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(B, -1), dim=-1).mean()
            step_size = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
            x_mean = x + step_size[:, None, None, None] * grad
            x = x_mean + torch.sqrt(step_size * 2)[:, None, None, None] * noise
        x = torch.where(cm, x, x_initial).float()
        # ---- reverse-diffusion predictor, sampling.py:162-167 + sde_lib.py:96-101,237-245 (f = 0 for VE)
        G = sde.discretize_G(vec_t)
        score = score_model(x, labels, context)
        rev_f = torch.zeros_like(x) - G[:, None, None, None] ** 2 * score * (0.5 if probability_flow else 1.0)
        rev_G = torch.zeros_like(G) if probability_flow else G
        z = noise_fn(stream_predictor(i, n_steps), x)
        x_mean = x - rev_f
        x = x_mean + rev_G[:, None, None, None] * z
        x = torch.where(cm, x, x_initial).float()
    x_mean = torch.where(cm, x_mean, x_initial).float()
    return (x_mean if denoise else x), K * (n_steps + 1)


def philox_noise_fn(seed, sample_offset=0):
    """noise_fn drawing from the numpy Philox restatement; ``sample_offset`` = global index of sample 0."""
    from .philox_ref import philox_normal

    def fn(stream, like):
        n = like.numel()
        per = n // like.shape[0]
        arr = philox_normal(seed, stream, sample_offset * per, n)
        return torch.from_numpy(arr).reshape(like.shape)

    return fn
