"""Generate the committed golden fixtures by running the UNMODIFIED reference (szhan227/text2protein) on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The GPU box never runs this; tests read the .npz / .json files written next to this script.

What is pinned:
  param_tree_<cfg>.json  state_dict names / shapes / order of ncsnpp.UNetModel   (ncsnpp.py:74-217)
  unet_tiny{5,8}.npz     score-net output + per-block activations, re-randomised weights (ncsnpp.py:220-263)
  tables.npz             sigma / timestep / label / G tables for N in {10,100,1000,2000} (sde_lib.py, models/utils.py)
  sampler_*.npz          full pc_sampler runs with injected Philox noise, with and without conditions
                         (sampling.py:245-289)
"""
import json
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("T2P_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle.philox_ref import STREAM_PRIOR, philox_normal, stream_corrector, stream_predictor  # noqa: E402
from oracle.unet_ref import AttrDict, rerandomize_  # noqa: E402
from tests.cfgs import synthetic_condition, synthetic_inputs, tiny_cfg  # noqa: E402

from score_sde_pytorch import sampling as ref_sampling  # noqa: E402
from score_sde_pytorch import sde_lib as ref_sde  # noqa: E402
from score_sde_pytorch.models import ncsnpp as ref_ncsnpp  # noqa: E402
from score_sde_pytorch.models import utils as ref_mutils  # noqa: E402


def build_ref_model(cfg, seed=42):
    torch.manual_seed(cfg.seed)
    model = ref_ncsnpp.UNetModel(cfg)
    rerandomize_(model.named_parameters(), seed)
    model.eval()
    return model


def param_tree(cfg, path):
    model = ref_ncsnpp.UNetModel(cfg)
    tree = {"state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()],
            "parameters": [k for k, _ in model.named_parameters()]}
    with open(path, "w") as f:
        json.dump(tree, f)
    print(path, len(tree["state_dict"]), "keys", sum(p.numel() for p in model.parameters()), "params")


def unet_golden(cfg, name, batch=2, ctx_len=8, keep_taps=True):
    model = build_ref_model(cfg)
    x, labels, ctx = synthetic_inputs(cfg, batch, ctx_len)
    taps = {}
    hooks = []

    def grab(key):
        def fn(_m, _i, o):
            taps[key] = o.detach().clone()
        return fn

    hooks.append(model.pre_conv.register_forward_hook(grab("pre_conv")))
    for i, blk in enumerate(model.input_blocks):
        hooks.append(blk.register_forward_hook(grab(f"input_blocks.{i}")))
    hooks.append(model.mid_blocks.register_forward_hook(grab("mid_blocks")))
    for i, blk in enumerate(model.out_blocks):
        hooks.append(blk.register_forward_hook(grab(f"out_blocks.{i}")))
    hooks.append(model.out[2].register_forward_hook(grab("out")))
    with torch.no_grad():
        out = model(x, labels, ctx)
        for h in hooks:
            h.remove()
        out_ctx2 = model(x, labels, ctx * 2.0)
        out_unit = model(x, labels, ctx / 0.02)  # unit-variance context: cross-attention far from uniform
    assert out.dtype == torch.float64
    sens = (out - out_ctx2).abs().max().item()
    print(name, "out absmax", out.abs().max().item(), "context sensitivity", sens)
    assert sens > 1e-6
    if not keep_taps:
        taps = {}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.numpy(), out_unit=out_unit.numpy(),
                        **{"tap:" + k: v.numpy() for k, v in taps.items()})


def tables():
    d = {}
    for N in (10, 100, 1000, 2000):
        sde = ref_sde.VESDE(sigma_min=0.01, sigma_max=100.0, N=N)
        ts = torch.linspace(sde.T, 1e-5, sde.N)
        cfg = AttrDict({"model": {"sigma_max": 100.0, "sigma_min": 0.01, "num_scales": N}})
        d[f"model_sigmas_{N}"] = ref_mutils.get_sigmas(cfg)
        d[f"discrete_sigmas_{N}"] = sde.discrete_sigmas.numpy()
        d[f"timesteps_{N}"] = ts.numpy()
        labels, tidx, G = [], [], []
        for i in range(N):
            t = torch.ones(1) * ts[i]
            lab = sde.T - t
            lab *= sde.N - 1
            labels.append(torch.round(lab).long().item())
            tidx.append((t * (sde.N - 1) / sde.T).long().item())
            G.append(sde.discretize(torch.zeros(1, 1, 1, 1), t)[1].item())
        d[f"labels_{N}"] = np.array(labels)
        d[f"tidx_{N}"] = np.array(tidx)
        d[f"G_{N}"] = np.array(G, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **d)
    print("tables ok")


def sampler_golden(cfg, name, kinds, batch=2, ctx_len=8, seed=2024):
    """Runs ref pc_sampler end to end (sde.N = cfg.model.num_scales iterations) with torch.randn_like and
    prior_sampling replaced by the Philox stream the CUDA kernel uses."""
    model = build_ref_model(cfg)
    _, _, ctx = synthetic_inputs(cfg, batch, ctx_len)
    cond = synthetic_condition(cfg, batch, kinds) if kinds else {}
    N = cfg.model.num_scales
    sde = ref_sde.VESDE(sigma_min=cfg.model.sigma_min, sigma_max=cfg.model.sigma_max, N=N)
    shape = (batch, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    n_steps = cfg.sampling.n_steps_each
    streams = []
    for i in range(N):
        streams += [stream_corrector(i, j, n_steps) for j in range(n_steps)] + [stream_predictor(i, n_steps)]
    it = iter(streams)
    numel = int(np.prod(shape))

    def fake_randn_like(x, *a, **k):
        return torch.from_numpy(philox_normal(seed, next(it), 0, numel)).reshape(shape)

    sde.prior_sampling = lambda shp: torch.from_numpy(philox_normal(seed, STREAM_PRIOR, 0, numel)).reshape(shape) \
        * sde.sigma_max
    fn = ref_sampling.get_sampling_fn(cfg, sde, shape, 1e-5)
    real = torch.randn_like
    torch.randn_like = fake_randn_like
    try:
        sample, nfe = fn(model, cond, ctx)
    finally:
        torch.randn_like = real
    assert next(it, None) is None
    assert sample.dtype == torch.float32
    np.savez_compressed(os.path.join(HERE, name + ".npz"), sample=sample.numpy(), nfe=np.array(nfe))
    print(name, "sample absmax", sample.abs().max().item(), "nfe", nfe)


def baseline_cfg(fname):
    with open(os.path.join(REF, "configs", fname)) as f:
        cfg = AttrDict(yaml.safe_load(f))
    if "n_heads" not in cfg.model:  # injected identically on both sides
        cfg.model.n_heads = 8
        cfg.model.context_dim = 4096
    cfg.device = "cpu"
    return cfg


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    param_tree(tiny_cfg(5), os.path.join(HERE, "param_tree_tiny5.json"))
    param_tree(baseline_cfg("cond_length.yml"), os.path.join(HERE, "param_tree_cond_length.json"))
    tables()
    unet_golden(tiny_cfg(5), "unet_tiny5")
    unet_golden(tiny_cfg(8), "unet_tiny8", keep_taps=False)
    sampler_golden(tiny_cfg(5), "sampler_tiny5_length", ["length"])
    sampler_golden(tiny_cfg(8), "sampler_tiny8_all", ["length", "ss", "inpainting"])
    sampler_golden(tiny_cfg(8), "sampler_tiny8_nocond", [])
