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
