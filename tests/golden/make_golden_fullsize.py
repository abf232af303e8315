"""BASELINE-size goldens from the UNMODIFIED reference (szhan227/text2protein) on CPU -- round 2.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_fullsize.py [names...]
The GPU box never runs this; tests read the files written next to this script.

What is pinned here, on top of make_golden.py's tiny-network fixtures:
  param_tree_<cfg>.json      state_dict names / shapes / order for all five BASELINE configs (ncsnpp.py:74-217)
  unet_full_<case>.npz       score-net output (float32 copy of the float64 result) of the real architectures at
                             B = 1 with re-randomised weights: cfg2 at L = 77 and L = 256, cfg3 (C = 8; no_cond.yml
                             builds the same network), test_config as shipped, cfg4 (N = 256, d_head = 128, L = 512)
                             + checksums of the inputs and a few parameters so that a drift of the torch generator
                             between the two machines is detected, not mis-read as a kernel bug
  sampler_full_cond_length.npz   K = 2 iterations of the reference pc_sampler at cfg2, B = 2, length condition
  sampler_vpsde_tiny5.npz    the reference pc_sampler with VPSDE (sde_lib.py:106-157, models/utils.py:139-156,
                             sampling.py:184-186) on the tiny network, length condition
  rsde.npz                   RSDE.sde / RSDE.discretize of VESDE and VPSDE, with and without probability_flow
                             (sde_lib.py:66-103), for an analytic score function
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import make_golden as mg  # noqa: E402  (puts /root/reference on sys.path and imports the reference modules)
from oracle.philox_ref import STREAM_PRIOR, philox_normal, stream_corrector, stream_generic, stream_predictor  # noqa: E402
from tests.cfgs import (FULLSIZE_CASES, analytic_score, fullsize_inputs, rsde_inputs, synthetic_condition,  # noqa: E402
                        synthetic_inputs, tiny_cfg)

ref_sampling, ref_sde, ref_ncsnpp = mg.ref_sampling, mg.ref_sde, mg.ref_ncsnpp


def weight_probe(model):
    """A few parameter checksums (first / middle / last tensors) of the re-randomised weights."""
    named = list(model.named_parameters())
    pick = [named[0], named[len(named) // 2], named[-1]]
    return np.array([p.double().sum().item() for _, p in pick])


def unet_full(case):
    fname, B, L = FULLSIZE_CASES[case]
    cfg = mg.baseline_cfg(fname)
    model = mg.build_ref_model(cfg)
    x, labels, ctx = fullsize_inputs(cfg, B, L)
    with torch.no_grad():
        out = model(x, labels, ctx)
    assert out.dtype == torch.float64 and torch.isfinite(out).all()
    np.savez_compressed(os.path.join(HERE, f"unet_full_{case}.npz"), out=out.float().numpy(),
                        in_sums=np.array([x.double().sum().item(), ctx.double().sum().item(),
                                          float(labels.sum().item())]),
                        w_sums=weight_probe(model), absmax=np.array(out.abs().max().item()))
    print(case, "out absmax", out.abs().max().item(), "params", sum(p.numel() for p in model.parameters()))


def truncated_tqdm(K):
    """The reference loop is ``for i in tqdm(range(sde.N))`` (sampling.py:279): the progress-bar wrapper is the
    one place a run can be cut to its first K iterations without touching the reference."""
    return lambda it, **kw: itertools.islice(it, K)


def sampler_full(K=2, B=2, L=77, seed=2024):
    cfg = mg.baseline_cfg("cond_length.yml")
    model = mg.build_ref_model(cfg)
    _, _, ctx = fullsize_inputs(cfg, B, L)
    cond = synthetic_condition(cfg, B, ["length"])
    N = cfg.model.num_scales
    sde = ref_sde.VESDE(sigma_min=cfg.model.sigma_min, sigma_max=cfg.model.sigma_max, N=N)
    shape = (B, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    n_steps = cfg.sampling.n_steps_each
    streams = []
    for i in range(K):
        streams += [stream_corrector(i, j, n_steps) for j in range(n_steps)] + [stream_predictor(i, n_steps)]
    it = iter(streams)
    numel = int(np.prod(shape))
    sde.prior_sampling = lambda shp: torch.from_numpy(philox_normal(seed, STREAM_PRIOR, 0, numel)).reshape(shape) \
        * sde.sigma_max
    fn = ref_sampling.get_sampling_fn(cfg, sde, shape, 1e-5)
    real, real_tqdm = torch.randn_like, ref_sampling.tqdm
    torch.randn_like = lambda x, *a, **k: torch.from_numpy(philox_normal(seed, next(it), 0, numel)).reshape(shape)
    ref_sampling.tqdm = truncated_tqdm(K)
    try:
        sample, _ = fn(model, cond, ctx)
    finally:
        torch.randn_like, ref_sampling.tqdm = real, real_tqdm
    assert next(it, None) is None and sample.dtype == torch.float32
    np.savez_compressed(os.path.join(HERE, "sampler_full_cond_length.npz"), sample=sample.numpy(), K=np.array(K),
                        w_sums=weight_probe(model))
    print("sampler_full_cond_length absmax", sample.abs().max().item())


VP_SCALES, VP_ITERS = 24, 4  # beta_max / N must stay below 1 (alpha = 1 - beta_max / N > 0); first 4 iterations


def sampler_vpsde(seed=2024, B=2, L=8):
    """VPSDE through the reference sampler: get_score_fn's VP branch (labels = t (N - 1) as FLOAT time
    conditioning, score = -out / sqrt(1 - alpha_bar)), the DDPM discretisation and the alpha-scaled Langevin step."""
    cfg = tiny_cfg(5, num_scales=VP_SCALES)
    model = mg.build_ref_model(cfg)
    _, _, ctx = synthetic_inputs(cfg, B, L)
    cond = synthetic_condition(cfg, B, ["length"])
    N = cfg.model.num_scales
    sde = ref_sde.VPSDE(beta_min=cfg.model.beta_min, beta_max=cfg.model.beta_max, N=N)
    shape = (B, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    numel = int(np.prod(shape))
    counter = itertools.count()
    sde.prior_sampling = lambda shp: torch.from_numpy(philox_normal(seed, STREAM_PRIOR, 0, numel)).reshape(shape)
    fn = ref_sampling.get_sampling_fn(cfg, sde, shape, 1e-3)
    real, real_tqdm = torch.randn_like, ref_sampling.tqdm
    torch.randn_like = lambda x, *a, **k: torch.from_numpy(
        philox_normal(seed, stream_generic(next(counter)), 0, numel)).reshape(shape)
    ref_sampling.tqdm = truncated_tqdm(VP_ITERS)
    try:
        sample, nfe = fn(model, cond, ctx)
    finally:
        torch.randn_like, ref_sampling.tqdm = real, real_tqdm
    assert next(counter) == 2 * VP_ITERS and sample.dtype == torch.float32 and torch.isfinite(sample).all()
    np.savez_compressed(os.path.join(HERE, "sampler_vpsde_tiny5.npz"), sample=sample.numpy(),
                        K=np.array(VP_ITERS), N=np.array(N))
    print("sampler_vpsde_tiny5 absmax", sample.abs().max().item(), "nfe", nfe)


def rsde():
    x, t = rsde_inputs()
    d = {}
    for name, sde in (("ve", ref_sde.VESDE(0.01, 100.0, 50)), ("vp", ref_sde.VPSDE(0.1, 20.0, 50))):
        for pf in (False, True):
            r = sde.reverse(analytic_score, probability_flow=pf)
            drift, diff = r.sde(x, t)
            f, G = r.discretize(x, t)
            tag = f"{name}_pf{int(pf)}"
            d[tag + "_drift"] = drift.numpy()
            d[tag + "_diffusion"] = np.asarray(diff if not torch.is_tensor(diff) else diff.numpy(), dtype=np.float64)
            d[tag + "_f"] = f.numpy()
            d[tag + "_G"] = G.numpy()
            assert r.N == 50 and r.T == 1
        fdrift, fdiff = sde.sde(x, t)
        mean, std = sde.marginal_prob(x, t)
        d[name + "_fwd_drift"], d[name + "_fwd_diffusion"] = fdrift.numpy(), fdiff.numpy()
        d[name + "_mean"], d[name + "_std"] = mean.numpy(), std.numpy()
        d[name + "_prior_logp"] = sde.prior_logp(x).numpy()
    np.savez_compressed(os.path.join(HERE, "rsde.npz"), **d)
    print("rsde ok", len(d), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    want = sys.argv[1:]

    def on(name):
        return not want or name in want

    if on("trees"):
        for fname in ("cond_ss_inpainting.yml", "no_cond.yml", "test_config.yml", "test_config_large.yml"):
            mg.param_tree(mg.baseline_cfg(fname), os.path.join(HERE, "param_tree_" + fname[:-4] + ".json"))
    if on("rsde"):
        rsde()
    if on("vpsde"):
        sampler_vpsde()
    for case in FULLSIZE_CASES:
        if on(case) or on("unet"):
            unet_full(case)
    if on("sampler"):
        sampler_full()
