"""Score-network parity: native engine (through UNetModel.forward -> C ABI) vs the oracle and the golden
outputs of the reference.  Tolerances are the ones BASELINE.json's north_star states: 1e-5 relative in fp32,
2e-2 in bf16 (relative = max |a - b| / max |b| over the tensor)."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_ref
from tests.cfgs import synthetic_inputs, tiny_cfg
from tests.gpu_util import make_native, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.mark.parametrize("c", [5, 8])
def test_fp32_matches_golden_reference_output(golden_dir, c):
    cfg, model, sd = make_native(tiny_cfg(c), "fp32")
    g = np.load(os.path.join(golden_dir, f"unet_tiny{c}.npz"))
    x, labels, ctx = synthetic_inputs(cfg, 2, 8)
    model.set_debug(True)
    out = model(x.cuda(), labels.cuda(), ctx.cuda())
    assert out.dtype == torch.float64 and out.shape == x.shape
    worst = []
    for k in g.files:
        if k.startswith("tap:"):
            worst.append((rel_err(model.tap(k[4:]), torch.from_numpy(g[k])), k))
    msg = " ".join(f"{k}={e:.1e}" for e, k in worst)
    assert rel_err(out, torch.from_numpy(g["out"])) < FP32_TOL, msg
    assert all(e < FP32_TOL * 3 for e, _ in worst), msg
    out_unit = model(x.cuda(), labels.cuda(), (ctx / 0.02).cuda())
    assert rel_err(out_unit, torch.from_numpy(g["out_unit"])) < FP32_TOL


@pytest.mark.parametrize("c", [5, 8])
def test_bf16_matches_golden_reference_output(golden_dir, c):
    cfg, model, sd = make_native(tiny_cfg(c), "bf16")
    g = np.load(os.path.join(golden_dir, f"unet_tiny{c}.npz"))
    x, labels, ctx = synthetic_inputs(cfg, 2, 8)
    out = model(x.cuda(), labels.cuda(), ctx.cuda())
    assert rel_err(out, torch.from_numpy(g["out"])) < BF16_TOL
    out_unit = model(x.cuda(), labels.cuda(), (ctx / 0.02).cuda())
    assert rel_err(out_unit, torch.from_numpy(g["out_unit"])) < BF16_TOL


@pytest.mark.parametrize("dtype,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_matches_oracle_on_other_shapes(dtype, tol):
    # deeper net: 4 levels, 2 blocks per level, attention at two resolutions, odd batch, ragged text length
    base = tiny_cfg(5, num_scales=50, max_res=64, nf=64, ch_mult=(1, 2, 2, 2), attn=(16, 8), n_heads=8,
                    context_dim=128, num_res_blocks=2)
    cfg, model, sd = make_native(base, dtype, seed=7)
    x, labels, ctx = synthetic_inputs(cfg, 3, 19, seed=5, ctx_scale=1.0)
    out = model(x.cuda(), labels.cuda(), ctx.cuda())
    ref = unet_ref.unet_forward(sd, cfg, x, labels, ctx)
    assert rel_err(out, ref) < tol


def test_weight_reload_and_context_change_are_picked_up():
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    x, labels, ctx = synthetic_inputs(cfg, 2, 8)
    out1 = model(x.cuda(), labels.cuda(), ctx.cuda())
    unet_ref.rerandomize_(model.named_parameters(), 123)       # e.g. ema.copy_to / load_state_dict
    sd2 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    out2 = model(x.cuda(), labels.cuda(), ctx.cuda())
    assert rel_err(out2, unet_ref.unet_forward(sd2, cfg, x, labels, ctx)) < FP32_TOL
    assert rel_err(out2, out1) > 1e-2
    ctx2 = ctx.flip(0).contiguous()
    out3 = model(x.cuda(), labels.cuda(), ctx2.cuda())
    assert rel_err(out3, unet_ref.unet_forward(sd2, cfg, x, labels, ctx2)) < FP32_TOL


def test_state_dict_roundtrip_through_dataparallel_checkpoint(tmp_path):
    from text2protein_b200.score_sde_pytorch import utils as sutils
    from text2protein_b200.score_sde_pytorch.models.ema import ExponentialMovingAverage
    from tests.cfgs import with_device
    cfg = with_device(tiny_cfg(5), "cuda")
    cfg.model.compute_dtype = "fp32"
    model = sutils.get_model(cfg)
    unet_ref.rerandomize_(model.named_parameters(), 3)
    ema = ExponentialMovingAverage(model.parameters(), decay=0.999)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    state = dict(optimizer=opt, model=model, ema=ema, step=5)
    path = str(tmp_path / "ckpt.pth")
    sutils.save_checkpoint(path, state)
    model2 = sutils.get_model(cfg)
    ema2 = ExponentialMovingAverage(model2.parameters(), decay=0.999)
    state2 = dict(optimizer=torch.optim.Adam(model2.parameters(), lr=1e-4), model=model2, ema=ema2, step=0)
    state2 = sutils.restore_checkpoint(path, state2, "cuda")
    ema2.copy_to(model2.parameters())
    x, labels, ctx = synthetic_inputs(cfg, 1, 8)
    a = model(x.cuda(), labels.cuda(), ctx.cuda())
    b = model2(x.cuda(), labels.cuda(), ctx.cuda())
    assert state2["step"] == 5 and torch.equal(a, b)


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_batch_slices_match_and_runs_are_deterministic(dtype, tol):
    """A batch agrees with its slices run separately (to rounding: the tile plan, and with it the grouping of the
    fp32 GroupNorm partial sums, depends on the launch size), and repeated runs are bit-identical."""
    cfg, model, sd = make_native(tiny_cfg(5), dtype)
    x, labels, ctx = synthetic_inputs(cfg, 8, 9, seed=11)
    x, labels, ctx = x.cuda(), labels.cuda(), ctx.cuda()
    whole = model(x, labels, ctx)
    again = model(x, labels, ctx)
    assert torch.equal(whole, again)
    lo = model(x[:2].contiguous(), labels[:2].contiguous(), ctx[:2].contiguous())   
    hi = model(x[2:].contiguous(), labels[2:].contiguous(), ctx[2:].contiguous())
    assert rel_err(whole, torch.cat([lo, hi])) < tol
    ref = unet_ref.unet_forward(sd, cfg, x.cpu(), labels.cpu(), ctx.cpu())
    assert rel_err(whole, ref) < (FP32_TOL if dtype == "fp32" else BF16_TOL)
