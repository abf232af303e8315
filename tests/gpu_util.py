"""Helpers shared by the GPU parity tests."""
import torch

from oracle import unet_ref
from tests.cfgs import tiny_cfg, with_device


def make_native(cfg, dtype, seed=42):
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
    cfg = with_device(cfg, "cuda")
    cfg.model.compute_dtype = dtype
    model = UNetModel(cfg).to("cuda")
    unet_ref.rerandomize_(model.named_parameters(), seed)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    return cfg, model, sd


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def nhwc(x, dtype=torch.float32):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()
