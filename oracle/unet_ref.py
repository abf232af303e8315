"""Oracle (test infrastructure): functional torch-CPU restatement of the reference score UNet.

Takes a plain ``state_dict`` (reference key names, no ``module.`` prefix) and an attribute-style config and
reproduces ``UNetModel.forward`` (score_sde_pytorch/models/ncsnpp.py:220-263) op for op.  Pinned against
outputs of the imported reference by tests/test_oracle_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


class AttrDict(dict):
    """Minimal stand-in for easydict.EasyDict (recursive attribute access; missing key -> AttributeError)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def get_sigmas(cfg):
    # score_sde_pytorch/models/utils.py:50-60 -- float64, DESCENDING (sigma_max first)
    m = cfg.model
    return np.exp(np.linspace(np.log(m.sigma_max), np.log(m.sigma_min), m.num_scales))


def timestep_embedding(labels, dim, max_positions=10000):
    # score_sde_pytorch/models/layers.py:97-111
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    arg = labels.float()[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def _gn(sd, key, x, groups=None):
    c = x.shape[1]
    g = min(c // 4, 32) if groups is None else groups
    return F.group_norm(x, g, sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _nin(sd, key, x):
    # layers.py:128-137 -- W is [in, out]
    y = torch.einsum("bchw,co->bohw", x, sd[key + ".W"])
    return y + sd[key + ".b"][None, :, None, None]


def _up(x):
    # layers.py:179-183 nearest x2 by repeat
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _down(x):
    # layers.py:185-188 2x2 mean
    b, c, h, w = x.shape
    return x.reshape(b, c, h // 2, 2, w // 2, 2).mean(dim=(3, 5))


def resblock(sd, key, x, temb, up=False, down=False):
    # ResnetBlockBigGANpp.forward, layers.py:303-327 (skip_rescale=True in every config)
    in_ch = x.shape[1]
    out_ch = sd[key + ".Conv_0.weight"].shape[0]
    h = F.silu(_gn(sd, key + ".GroupNorm_0", x))
    if up:
        h, x = _up(h), _up(x)
    elif down:
        h, x = _down(h), _down(x)
    h = F.conv2d(h, sd[key + ".Conv_0.weight"], sd[key + ".Conv_0.bias"], padding=1)
    h = h + F.linear(F.silu(temb), sd[key + ".Dense_0.weight"], sd[key + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(sd, key + ".GroupNorm_1", h))
    h = F.conv2d(h, sd[key + ".Conv_1.weight"], sd[key + ".Conv_1.bias"], padding=1)
    if in_ch != out_ch or up or down:
        x = F.conv2d(x, sd[key + ".Conv_2.weight"], sd[key + ".Conv_2.bias"])
    return (x + h) / np.sqrt(2.0)


def attnblock(sd, key, x):
    # AttnBlockpp.forward, layers.py:160-176
    b, c, hh, ww = x.shape
    h = _gn(sd, key + ".GroupNorm_0", x)
    q = _nin(sd, key + ".NIN_0", h).reshape(b, c, hh * ww)
    k = _nin(sd, key + ".NIN_1", h).reshape(b, c, hh * ww)
    v = _nin(sd, key + ".NIN_2", h).reshape(b, c, hh * ww)
    w = torch.einsum("bct,bcs->bts", q, k) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=-1)
    h = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, hh, ww)
    h = _nin(sd, key + ".NIN_3", h)
    return (x + h) / np.sqrt(2.0)


def cross_attention(sd, key, x, context, heads):
    # CrossAttention.forward, model/attention.py:170-193 (no mask is ever passed on this path)
    q = F.linear(x, sd[key + ".to_q.weight"])
    ctx = x if context is None else context
    k = F.linear(ctx, sd[key + ".to_k.weight"])
    v = F.linear(ctx, sd[key + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, n, inner)
    return F.linear(out, sd[key + ".to_out.0.weight"], sd[key + ".to_out.0.bias"])


def spatial_transformer(sd, key, x, context, heads):
    # SpatialTransformer.forward, model/attention.py:250-263; block :211-215; GEGLU :42-44
    b, c, hh, ww = x.shape
    x_in = x
    h = _gn(sd, key + ".norm", x, groups=32)
    h = F.conv2d(h, sd[key + ".proj_in.weight"], sd[key + ".proj_in.bias"])
    t = h.reshape(b, c, hh * ww).permute(0, 2, 1)
    bk = key + ".transformer_blocks.0"
    dim = t.shape[-1]

    def ln(name, z):
        return F.layer_norm(z, (dim,), sd[f"{bk}.{name}.weight"], sd[f"{bk}.{name}.bias"])

    t = cross_attention(sd, bk + ".attn1", ln("norm1", t), None, heads) + t
    t = cross_attention(sd, bk + ".attn2", ln("norm2", t), context, heads) + t
    z = F.linear(ln("norm3", t), sd[bk + ".ff.net.0.proj.weight"], sd[bk + ".ff.net.0.proj.bias"])
    a, gate = z.chunk(2, dim=-1)
    z = a * F.gelu(gate)
    t = F.linear(z, sd[bk + ".ff.net.2.weight"], sd[bk + ".ff.net.2.bias"]) + t
    h = t.permute(0, 2, 1).reshape(b, c, hh, ww)
    h = F.conv2d(h, sd[key + ".proj_out.weight"], sd[key + ".proj_out.bias"])
    return h + x_in


def block_plan(cfg):
    """Module layout of UNetModel.__init__ (ncsnpp.py:141-208): lists of (kind, flags) per block."""
    m = cfg.model
    nres = len(m.ch_mult)
    res = [cfg.data.max_res_num // (2 ** i) for i in range(nres)]
    inp, out = [], []
    for lvl in range(nres):
        for _ in range(m.num_res_blocks):
            mods = [("res", {})]
            if res[lvl] in m.attn_resolutions:
                mods += [("attn", {}), ("st", {})]
            inp.append(mods)
        if lvl != nres - 1:
            inp.append([("res", {"down": True})])
    mid = [("res", {}), ("attn", {}), ("st", {}), ("res", {})]
    for lvl in reversed(range(nres)):
        for ib in range(m.num_res_blocks + 1):
            mods = [("res", {})]
            if res[lvl] in m.attn_resolutions:
                mods += [("attn", {}), ("st", {})]
            if lvl != 0 and ib == m.num_res_blocks:
                mods.append(("res", {"up": True}))
            out.append(mods)
    return inp, mid, out


def _run_seq(sd, prefix, mods, h, temb, ctx, heads):
    # TimestepEmbedSequential.forward, ncsnpp.py:54-69
    for j, (kind, flags) in enumerate(mods):
        key = f"{prefix}.{j}"
        if kind == "res":
            h = resblock(sd, key, h, temb, **flags)
        elif kind == "attn":
            h = attnblock(sd, key, h)
        else:
            h = spatial_transformer(sd, key, h, ctx, heads)
    return h


@torch.no_grad()
def unet_forward(sd, cfg, x, labels, context, taps=None):
    """Returns float64 [B,C,N,N] exactly like the reference (h / sigmas promotes).

    ``taps`` (optional dict) receives the fp32 output of pre_conv / every input, mid and out block and the
    un-scaled final conv under the names the native engine uses for its debug taps.
    """
    m = cfg.model
    assert m.resblock_type.lower() == "biggan" and m.embedding_type.lower() == "positional"
    assert m.nonlinearity.lower() == "swish" and m.skip_rescale
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
    heads = m.n_heads
    inp, mid, out = block_plan(cfg)
    sigmas = sd["sigmas"] if "sigmas" in sd else torch.tensor(get_sigmas(cfg))
    used_sigmas = sigmas[labels.long()]
    temb = timestep_embedding(labels, m.nf)
    # NOTE no activation between the two pre_blocks Linears (ncsnpp.py:227-228)
    temb = F.linear(temb, sd["pre_blocks.0.weight"], sd["pre_blocks.0.bias"])
    temb = F.linear(temb, sd["pre_blocks.1.weight"], sd["pre_blocks.1.bias"])
    h = F.conv2d(x.float(), sd["pre_conv.weight"], sd["pre_conv.bias"], padding=1)
    if taps is not None:
        taps["pre_conv"] = h
    hs = [h]
    for i, mods in enumerate(inp):
        h = _run_seq(sd, f"input_blocks.{i}", mods, h, temb, context, heads)
        hs.append(h)
        if taps is not None:
            taps[f"input_blocks.{i}"] = h
    h = _run_seq(sd, "mid_blocks", mid, h, temb, context, heads)
    if taps is not None:
        taps["mid_blocks"] = h
    for i, mods in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_seq(sd, f"out_blocks.{i}", mods, h, temb, context, heads)
        if taps is not None:
            taps[f"out_blocks.{i}"] = h
    assert not hs
    h = F.silu(_gn(sd, "out.0", h))
    h = F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    if taps is not None:
        taps["out"] = h
    if m.scale_by_sigma:
        h = h / used_sigmas.reshape(-1, 1, 1, 1)
    return h


def rerandomize_(named_tensors, seed):
    """Weight re-randomisation recipe of SURVEY.md 8(c)(ii): the as-shipped init (init_scale 0, zeroed
    proj_out) gives an output with zero context sensitivity, useless for parity.  Applied in
    ``named_parameters()`` order with one generator so any module tree with the reference's names, shapes and
    order receives identical values."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in named_tensors:
            leaf = name.rsplit(".", 1)[-1]
            parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
            is_norm = parent.startswith("GroupNorm") or parent.startswith("norm") or name.startswith("out.0") \
                or name.startswith("module.out.0")
            if p.dim() > 1:
                if leaf == "W":  # NIN: [in, out]
                    fan_in = p.shape[0]
                else:
                    fan_in = int(np.prod(p.shape[1:]))
                p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
            elif is_norm and leaf == "weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def state_dict_from_tree(tree, cfg, seed):
    """Builds a re-randomised state_dict from a committed ``param_tree_*.json`` (names, shapes and order of the
    reference module), so the oracle can run without the reference or the product module."""
    shapes = {k: s for k, s, _ in tree["state_dict"]}
    params = [(k, torch.empty(shapes[k])) for k in tree["parameters"]]
    rerandomize_(params, seed)
    sd = {"sigmas": torch.tensor(get_sigmas(cfg))}
    sd.update(dict(params))
    return sd
