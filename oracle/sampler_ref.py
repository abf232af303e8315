"""Oracle (test infrastructure): torch-CPU restatement of the reverse-SDE predictor-corrector sampler.

Follows score_sde_pytorch/sampling.py:157-289 and sde_lib.py:106-157,199-245 including the dtype promotions
of the reference (the score arrives as float64, the update runs in float64 and is cast back with
``.float()`` after each mask application).  Noise is injected through ``noise_fn(stream, like)`` so that a
run can be replayed with the exact normals the CUDA kernel generated.  Pinned against the imported reference
by tests/test_oracle_golden.py (VE runs with and without conditions, a VP run, and a K = 2 run of the real
cond_length.yml network at N = 128).
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

    def timestep(self, t):
        return (t * (self.N - 1) / self.T).long()


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


def symmetrize_free(u, cm):
    """The opt-in symmetrisation of the step kernels (NO reference counterpart; this function is its
    specification): channels 0 and 1 of ``u`` take 0.5 (u[i][j] + u[j][i]) wherever (i, j) and (j, i) are both
    free in ``cm``; everything else is left alone."""
    out = u.clone()
    both = cm[:, :2] & cm[:, :2].transpose(2, 3)
    sym = 0.5 * (u[:, :2] + u[:, :2].transpose(2, 3))
    out[:, :2] = torch.where(both, sym, u[:, :2])
    return out


@torch.no_grad()
def pc_sampler_ref(sde, score_model, shape, snr, n_steps=1, probability_flow=False, denoise=True, eps=1e-5,
                   condition=None, context=None, noise_fn=None, num_iters=None, x0=None, symmetrize=False,
                   generic_streams=False):
    """``score_model(x, labels, context) -> float64 [B,C,N,N]`` (e.g. oracle.unet_ref.unet_forward bound to a
    state_dict).  ``noise_fn(stream, like) -> tensor`` supplies every normal draw; ``num_iters`` truncates the
    loop to the first K of sde.N iterations (bench / parity at small K).  VESDERef (every shipped config) or
    VPSDERef (sde_lib.py:106-157 + the VP branches of models/utils.py:139-156 and sampling.py:184-186).
    ``generic_streams``: noise stream ids of the product's generic update_fn path (philox_ref.stream_generic)
    instead of the native loop's.  ``symmetrize``: see symmetrize_free (product extension, default off)."""
    from .philox_ref import STREAM_PRIOR, stream_corrector, stream_generic, stream_predictor

    vp = isinstance(sde, VPSDERef)
    assert vp or isinstance(sde, VESDERef)
    B = shape[0]
    if x0 is None:
        x = noise_fn(STREAM_PRIOR, torch.empty(shape)) * (1.0 if vp else sde.sigma_max)  # sde_lib.py:136-137,229-230
    else:
        x = x0.clone()
    timesteps = torch.linspace(sde.T, eps, sde.N)
    x, cm = build_condition(x, condition)
    x_initial = x.detach().clone()
    x_mean = x
    K = sde.N if num_iters is None else num_iters
    draws = 0

    def score_fn(x, vec_t):
        if vp:  # models/utils.py:151-156: float time conditioning, score = -out / sqrt(1 - alpha_bar)
            labels = vec_t * (sde.N - 1)
            out = score_model(x, labels, context)
            std = sde.sqrt_1m_alphas_cumprod[labels.long()]
            return -out / std[:, None, None, None]
        return score_model(x, sde.labels(vec_t.clone()), context)

    for i in range(K):
        vec_t = torch.ones(B) * timesteps[i]
        # ---- Langevin corrector, sampling.py:179-199 (alpha = 1 for VE)
        alpha = sde.alphas[sde.timestep(vec_t)] if vp else torch.ones_like(vec_t)
        for j in range(n_steps):
            grad = score_fn(x, vec_t)
            noise = noise_fn(stream_generic(draws) if generic_streams else stream_corrector(i, j, n_steps), x)
            draws += 1
            grad_norm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
            noise_norm = torch.norm(noise.reshape(B, -1), dim=-1).mean()
            step_size = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
            x_mean = x + step_size[:, None, None, None] * grad
            x = x_mean + torch.sqrt(step_size * 2)[:, None, None, None] * noise
            if symmetrize:  # inner steps before the last run unconditioned (the mask is applied after the corrector)
                free = cm if j == n_steps - 1 else torch.ones_like(cm)
                x, x_mean = symmetrize_free(x, free), symmetrize_free(x_mean, free)
        x = torch.where(cm, x, x_initial).float()
        # ---- reverse-diffusion predictor, sampling.py:162-167 + sde_lib.py:96-101 over :149-157 (VP) / :237-245 (VE)
        score = score_fn(x, vec_t)
        if vp:
            ts = sde.timestep(vec_t)
            f = torch.sqrt(sde.alphas[ts])[:, None, None, None] * x - x
            G = torch.sqrt(sde.discrete_betas[ts])
        else:
            f = torch.zeros_like(x)
            G = sde.discretize_G(vec_t)
        rev_f = f - G[:, None, None, None] ** 2 * score * (0.5 if probability_flow else 1.0)
        rev_G = torch.zeros_like(G) if probability_flow else G
        z = noise_fn(stream_generic(draws) if generic_streams else stream_predictor(i, n_steps), x)
        draws += 1
        x_mean = x - rev_f
        x = x_mean + rev_G[:, None, None, None] * z
        if symmetrize:
            x, x_mean = symmetrize_free(x, cm), symmetrize_free(x_mean, cm)
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
