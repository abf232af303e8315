"""Synthetic weights / inputs for benchmarks and smoke runs (there are no checkpoints or datasets offline).

``rerandomize_`` is the weight recipe of SURVEY.md 8(c)(ii): the as-shipped random init (init_scale 0, zeroed
proj_out) produces an output that ignores the text context, so benchmarks and parity runs redraw every parameter.
tests/test_abi.py checks that it draws exactly the values of the oracle's copy.
"""
import numpy as np
import torch


def rerandomize_(named_tensors, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in named_tensors:
            parts = name.split(".")
            leaf = parts[-1]
            parent = parts[-2] if len(parts) > 1 else ""
            is_norm = parent.startswith("GroupNorm") or parent.startswith("norm") or ".".join(parts[-3:-1]) == "out.0" \
                or name.startswith("out.0.")
            if p.dim() > 1:
                fan_in = p.shape[0] if leaf == "W" else int(np.prod(p.shape[1:]))
                p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
            elif is_norm and leaf == "weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def rerandomize_device_(named_tensors, seed):
    """Same recipe drawn with the DEVICE generator of each tensor (timing runs of the large configurations: no
    3.4 GB of host random numbers; the values differ from ``rerandomize_``'s, the distribution does not)."""
    gens = {}
    with torch.no_grad():
        for name, p in named_tensors:
            g = gens.get(p.device)
            if g is None:
                g = gens[p.device] = torch.Generator(device=p.device).manual_seed(seed)
            parts = name.split(".")
            leaf = parts[-1]
            parent = parts[-2] if len(parts) > 1 else ""
            is_norm = parent.startswith("GroupNorm") or parent.startswith("norm") or name.startswith("out.0.")
            r = torch.randn(p.shape, generator=g, device=p.device)
            if p.dim() > 1:
                fan_in = p.shape[0] if leaf == "W" else int(np.prod(p.shape[1:]))
                p.copy_(r * fan_in ** -0.5)
            elif is_norm and leaf == "weight":
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.1 * r)
