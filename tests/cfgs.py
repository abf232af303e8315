"""Shared test configurations and synthetic-input builders (used by the golden generator and the tests)."""
import copy

import torch

from oracle.unet_ref import AttrDict


def tiny_cfg(channels=5, num_scales=4, max_res=32, nf=64, ch_mult=(1, 1, 2), attn=(16,), n_heads=4,
             context_dim=64, num_res_blocks=1):
    """A few-hundred-thousand-parameter UNet with every module kind of the real one: down/up ResBlocks, skip
    concat with a channel change, AttnBlockpp + SpatialTransformer on a level and in the mid block."""
    return AttrDict({
        "training": {"sde": "vesde"},
        "sampling": {"n_steps_each": 1, "noise_removal": True, "probability_flow": False, "snr": 0.17,
                     "method": "pc", "predictor": "reverse_diffusion", "corrector": "langevin"},
        "data": {"max_res_num": max_res, "min_res_num": 8, "num_channels": channels},
        "model": {"condition": [], "sigma_max": 100.0, "sigma_min": 0.01, "num_scales": num_scales,
                  "beta_min": 0.1, "beta_max": 20.0, "dropout": 0.1, "embedding_type": "positional",
                  "name": "ncsnpp", "scale_by_sigma": True, "ema_rate": 0.999, "normalization": "GroupNorm",
                  "nonlinearity": "swish", "nf": nf, "ch_mult": list(ch_mult), "num_res_blocks": num_res_blocks,
                  "attn_resolutions": list(attn), "resamp_with_conv": True, "skip_rescale": True,
                  "resblock_type": "biggan", "attention_type": "ddpm", "init_scale": 0.0, "fourier_scale": 16,
                  "conv_size": 3, "n_heads": n_heads, "context_dim": context_dim},
        "seed": 42,
        "device": "cpu",
    })


def with_device(cfg, device):
    c = copy.deepcopy(cfg)
    c.device = device
    return c


def synthetic_inputs(cfg, batch, ctx_len, seed=1234, ctx_scale=0.02):
    g = torch.Generator().manual_seed(seed)
    C, N = cfg.data.num_channels, cfg.data.max_res_num
    x = torch.randn(batch, C, N, N, generator=g) * 3.0
    labels = torch.randint(0, cfg.model.num_scales, (batch,), generator=g)
    ctx = torch.randn(batch, ctx_len, cfg.model.context_dim, generator=g) * ctx_scale
    return x, labels, ctx


def synthetic_condition(cfg, batch, kinds, seed=77):
    """Self-consistent condition dict: coords_6d[:, -1] == length mask, coords_6d[:, 4:7] == ss."""
    g = torch.Generator().manual_seed(seed)
    C, N = cfg.data.num_channels, cfg.data.max_res_num
    cond = {}
    lengths = torch.randint(max(4, N // 3), N + 1, (batch,), generator=g)
    ar = torch.arange(N)
    lmask = (ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])
    if "length" in kinds:
        cond["length"] = lmask
    ss = None
    if "ss" in kinds:
        assert C >= 8
        blk = (torch.rand(batch, 3, N, generator=g) > 0.5).float()
        ss = blk[:, :, :, None] * blk[:, :, None, :] * lmask[:, None].float()
        cond["ss"] = ss
    if "inpainting" in kinds:
        coords = torch.rand(batch, C, N, N, generator=g) * 2 - 1
        coords[:, 0] = 0.5 * (coords[:, 0] + coords[:, 0].transpose(1, 2))
        coords[:, 1] = 0.5 * (coords[:, 1] + coords[:, 1].transpose(1, 2))
        coords = coords * lmask[:, None].float()
        if ss is not None:
            coords[:, 4:7] = ss
        coords[:, -1] = lmask.float()
        start = (lengths.float() * 0.2).long()
        stop = (lengths.float() * 0.7).long()
        r = (ar[None, :] >= start[:, None]) & (ar[None, :] < stop[:, None])
        cond["inpainting"] = {"coords_6d": coords, "mask_inpaint": r[:, :, None] | r[:, None, :]}
    return cond


def fullsize_inputs(cfg, batch, ctx_len, seed=3, ctx_scale=0.02):
    """CPU inputs of the BASELINE-size parity runs (tests/golden/make_golden_fullsize.py and
    tests/test_gpu_fullsize.py draw exactly these; the golden files carry checksums of them)."""
    g = torch.Generator().manual_seed(seed)
    C, N = cfg.data.num_channels, cfg.data.max_res_num
    x = torch.randn(batch, C, N, N, generator=g) * 5
    labels = torch.randint(0, cfg.model.num_scales, (batch,), generator=g)
    ctx = torch.randn(batch, ctx_len, cfg.model.context_dim, generator=g) * ctx_scale
    return x, labels, ctx


# BASELINE.json configurations pinned at full size: name -> (yaml, batch, context length)
FULLSIZE_CASES = {
    "cond_length_L77": ("cond_length.yml", 1, 77),
    "cond_length_L256": ("cond_length.yml", 1, 256),
    "cond_ss_inpainting": ("cond_ss_inpainting.yml", 1, 256),  # no_cond.yml builds the identical network (C = 8)
    "test_config": ("test_config.yml", 1, 32),                 # as shipped: N = 256, nf = 256, attention at 32/16/8
    "test_config_large": ("test_config_large.yml", 1, 512),    # N = 256, ch_mult [1,1,2,2,2,4], d_head = 128, L = 512
}


def rsde_inputs():
    """Inputs of the RSDE / SDE golden (tests/golden/rsde.npz)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 2, 4, 4, generator=g)
    t = torch.tensor([1.0, 0.43, 1e-3])
    return x, t


def analytic_score(x, t, context=None):
    return torch.sin(x) * (1.0 + t)[:, None, None, None]


def synthetic_backbone(nres, seed):
    """A self-avoiding-ish random-walk protein backbone: N, CA, C coordinates [nres, 3, 3] (Angstrom) with realistic
    bond lengths, compact enough that many C-beta pairs fall inside and many outside the 20 A cut-off."""
    import numpy as np
    rng = np.random.default_rng(seed)
    ca = np.zeros((nres, 3))
    d = rng.normal(size=3)
    for i in range(1, nres):
        d = d + 0.9 * rng.normal(size=3)
        d /= np.linalg.norm(d)
        ca[i] = ca[i - 1] + 3.8 * d
    xyz = np.zeros((nres, 3, 3))
    for i in range(nres):
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        v = rng.normal(size=3)
        v -= v.dot(u) * u
        v /= np.linalg.norm(v)
        xyz[i, 1] = ca[i]
        xyz[i, 0] = ca[i] + 1.46 * (-0.5 * u + 0.866 * v)   # N
        xyz[i, 2] = ca[i] + 1.52 * (-0.5 * u - 0.866 * v)   # C
    return xyz


def write_pdb(path, xyz, chain="A", missing=(), extra_chain=True):
    """Minimal PDB text of a backbone (fixed-column ATOM records), with optional missing atoms [(residue, atom name)],
    a second chain, an alternate location and a water HETATM -- everything the reader must skip."""
    names = ("N", "CA", "C")
    lines, serial = [], 1
    for i in range(xyz.shape[0]):
        for j, a in enumerate(names):
            if (i, a) in missing:
                continue
            x, y, z = xyz[i, j]
            alt = "A" if (i == 2 and a == "CA") else " "
            lines.append("ATOM  %5d %-4s%1s%3s %1s%4d    %8.3f%8.3f%8.3f  1.00  0.00" % (serial, " " + a, alt, "ALA", chain,
                                                                                       i + 1, x, y, z))
            serial += 1
            if alt == "A":  # second alternate location of the same atom: ignored by the reader (first one wins)
                lines.append("ATOM  %5d %-4s%1s%3s %1s%4d    %8.3f%8.3f%8.3f  0.50  0.00" % (serial, " " + a, "B", "ALA",
                                                                                           chain, i + 1, x + 5, y, z))
                serial += 1
        lines.append("ATOM  %5d %-4s%1s%3s %1s%4d    %8.3f%8.3f%8.3f  1.00  0.00" % (serial, " O", " ", "ALA", chain, i + 1,
                                                                                   0.0, 0.0, 0.0))
        serial += 1
    if extra_chain:
        lines.append("ATOM  %5d %-4s%1s%3s %1s%4d    %8.3f%8.3f%8.3f  1.00  0.00" % (serial, " CA", " ", "GLY", "B", 1, 1.0, 2.0, 3.0))
        lines.append("HETATM%5d %-4s%1s%3s %1s%4d    %8.3f%8.3f%8.3f  1.00  0.00" % (serial + 1, " O", " ", "HOH", chain, 900,
                                                                                   9.0, 9.0, 9.0))
    lines.append("END")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
