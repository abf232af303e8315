"""Condition builders of the reference's root ``utils.py`` (the callers just before the sampling loop), generated on
the device from per-sample integers instead of being assembled on the host and copied.

Same names, arguments and return values as the reference where the reference function is pure tensor logic:
``get_mask_all_lengths`` (utils.py:139-148), ``selected_mask_batch`` (utils.py:62-81) and the "length" / "ss" /
"inpainting" assembly of ``get_condition_from_batch`` (utils.py:83-106).  The dataset / PDB readers behind
``get_conditions_random`` and ``get_conditions_from_pdb`` (biotite, ProteinDataset) are out of scope.
"""
import ctypes as C

import torch

from . import _lib


def _dev(config):
    return torch.device(getattr(config, "device", "cuda"))


def parse_mask_info(mask_info):
    """'1:5,10:15,20' -> [[1, 5], [10, 15], [20, 20]] -- inclusive residue ranges, utils.py:70-77."""
    out = []
    for r in mask_info.split(","):
        if ":" in r:
            a, b = r.split(":")
            out.append([int(a), int(b)])
        else:
            out.append([int(r), int(r)])
    return out


def length_mask(lengths, max_res_num, device="cuda"):
    """bool [B, N, N], True inside the top-left len x len block (utils.py:89-93)."""
    lengths = torch.as_tensor(lengths, dtype=torch.int32, device=device).contiguous()
    B = lengths.numel()
    out = torch.empty(B, max_res_num, max_res_num, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().t2p_length_mask(_lib.ptr(lengths), B, max_res_num, _lib.ptr(out), _lib.current_stream()))
    return out.view(torch.bool)


def inpaint_mask(ranges, batch_size, max_res_num, device="cuda"):
    """bool [B, N, N]: rows and columns of the selected residues.  ``ranges`` is [[start, end], ...] (inclusive,
    shared by the batch, as ``selected_mask_batch`` applies ``mask_info``) or a [B, R, 2] tensor (per sample, as the
    outer-OR of ``random_mask_batch``, utils.py:56-58)."""
    r = torch.as_tensor(ranges, dtype=torch.int32, device=device).contiguous()
    per_sample = 1 if r.dim() == 3 else 0
    R = r.shape[-2] if r.numel() else 0
    out = torch.empty(batch_size, max_res_num, max_res_num, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().t2p_inpaint_mask(_lib.ptr(r) if R else C.c_void_p(0), R, per_sample, batch_size, max_res_num,
                                           _lib.ptr(out), _lib.current_stream()))
    return out.view(torch.bool)


def get_mask_all_lengths(config, batch_size=16):
    """bool [L, B, N, N] for every length in [min_res_num, max_res_num] (utils.py:139-148), built on the device."""
    N = config.data.max_res_num
    lens = torch.arange(config.data.min_res_num, N + 1, dtype=torch.int32)
    m = length_mask(lens, N, _dev(config))                       # [L, N, N]
    return m[:, None].expand(len(lens), batch_size, N, N).contiguous()


def selected_mask_batch(batch, mask_info, config):
    """utils.py:62-81: adds batch["mask_inpaint"] (bool [B, N, N]) for the residues named by ``mask_info``."""
    if "inpainting" not in config.model.condition:
        batch["mask_inpaint"] = None
        return batch
    B, _, N, _ = batch["coords_6d"].shape
    batch["mask_inpaint"] = inpaint_mask(parse_mask_info(mask_info), B, N, _dev(config))
    return batch


def get_condition_from_lengths(config, lengths, coords_6d=None, mask_info=None):
    """The condition dict ``get_condition_from_batch`` (utils.py:83-106) produces, from residue counts instead of
    ``batch["aa_str"]``: keys in ``config.model.condition`` order ("length", "ss", "inpainting")."""
    dev = _dev(config)
    out = {}
    for c in config.model.condition:
        if c == "length":
            out[c] = length_mask(lengths, config.data.max_res_num, dev)
        elif c == "ss":
            out[c] = coords_6d[:, 4:7].to(dev)
        elif c == "inpainting":
            B = coords_6d.shape[0]
            ranges = parse_mask_info(mask_info) if mask_info is not None else []
            out[c] = {"coords_6d": coords_6d.to(dev),
                      "mask_inpaint": inpaint_mask(ranges, B, config.data.max_res_num, dev)}
    return out


def conditional_mask(lengths, ranges, has_ss, batch_size, num_channels, max_res_num, device="cuda"):
    """The sampler's bool ``conditional_mask`` [B, C, N, N] (True = free to evolve, sampling.py:258-281) generated
    directly from integers; ``lengths`` / ``ranges`` may be None when that condition is absent."""
    l = torch.as_tensor(lengths, dtype=torch.int32, device=device).contiguous() if lengths is not None else None
    r = torch.as_tensor(ranges, dtype=torch.int32, device=device).contiguous() if ranges is not None else None
    per_sample = 1 if (r is not None and r.dim() == 3) else 0
    R = r.shape[-2] if r is not None else 0
    out = torch.empty(batch_size, num_channels, max_res_num, max_res_num, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().t2p_condition_mask(_lib.ptr(l), _lib.ptr(r), R, per_sample, int(bool(has_ss)), batch_size,
                                             num_channels, max_res_num, _lib.ptr(out), _lib.current_stream()))
    return out.view(torch.bool)


def embed_tokens(table, tokens):
    """``llm.model.embed_tokens(tokens)`` (sampling_6d.py:134-137) as a device gather: fp32 [B, L, D]."""
    assert table.is_cuda and table.dim() == 2 and table.dtype in (torch.float32, torch.bfloat16)
    tok = tokens.to(device=table.device, dtype=torch.int64).contiguous()
    out = torch.empty(*tok.shape, table.shape[1], dtype=torch.float32, device=table.device)
    _lib.check(_lib.lib().t2p_embed_tokens(_lib.ptr(table.contiguous()), _lib.torch_dtype_code(table.dtype),
                                           table.shape[0], table.shape[1], _lib.ptr(tok), tok.numel(), _lib.ptr(out),
                                           _lib.current_stream()))
    return out
