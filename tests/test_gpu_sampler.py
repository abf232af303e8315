"""PC-sampler parity: the native loop (t2p_pc_run, graph-replayed) and the generic update_fn path vs the oracle
fed the identical Philox normals; mask / index handling bit-exact; golden runs of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_ref, unet_ref
from tests.cfgs import synthetic_condition, synthetic_inputs, tiny_cfg
from tests.gpu_util import make_native, rel_err

pytestmark = pytest.mark.gpu


def _to_dev(cond):
    out = {}
    for k, v in cond.items():
        out[k] = {a: b.cuda() for a, b in v.items()} if isinstance(v, dict) else v.cuda()
    return out


def _run_native(cfg, model, ctx, cond, seed, **kw):
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    shape = (ctx.shape[0], cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    fn = sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                 snr=cfg.sampling.snr, n_steps=cfg.sampling.n_steps_each, eps=1e-5, device="cuda",
                                 seed=seed, **kw)
    s, nfe = fn(model, _to_dev(cond), ctx.cuda())
    torch.cuda.synchronize()
    return s.cpu(), nfe


def _run_oracle(cfg, sd, ctx, cond, seed, num_iters=None, gpu_noise=True):
    from text2protein_b200.score_sde_pytorch import sampling
    sde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    shape = (ctx.shape[0], cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    if gpu_noise:
        def noise_fn(stream, like):
            return sampling.philox_normal(tuple(like.shape), seed, stream, "cuda").cpu()
    else:
        noise_fn = sampler_ref.philox_noise_fn(seed)
    return sampler_ref.pc_sampler_ref(sde, lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), shape,
                                      cfg.sampling.snr, n_steps=cfg.sampling.n_steps_each, eps=1e-5, condition=cond,
                                      context=ctx, noise_fn=noise_fn, num_iters=num_iters)


def _check_masks(sample, cond):
    if "length" in cond:
        assert torch.equal(sample[:, -1], cond["length"].float())
    if "ss" in cond:
        assert torch.equal(sample[:, 4:7], cond["ss"])
    if "inpainting" in cond:
        keep = ~cond["inpainting"]["mask_inpaint"][:, None].expand_as(sample)
        assert torch.equal(sample[keep], cond["inpainting"]["coords_6d"][keep])


# K = 4 corrector+predictor iterations; after them |x| ~ 1e2..1e3, so 1e-4 relative = a few 1e-2 absolute.  bf16: the
# stated tolerance of the score network itself (2e-2); the measured deviation of the K-step maps is ~2e-3.
@pytest.mark.parametrize("c,kinds", [(5, ["length"]), (8, ["length", "ss", "inpainting"]), (8, [])])
@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_native_loop_matches_oracle(c, kinds, dtype, tol):
    cfg, model, sd = make_native(tiny_cfg(c), dtype)
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, kinds) if kinds else {}
    sample, nfe = _run_native(cfg, model, ctx, cond, seed=2024)
    ref, ref_nfe = _run_oracle(cfg, sd, ctx, cond, seed=2024)
    assert nfe == ref_nfe == 8 and sample.dtype == torch.float32
    _check_masks(sample, cond)
    assert rel_err(sample, ref) < tol


@pytest.mark.parametrize("name,c,kinds", [("sampler_tiny5_length", 5, ["length"]),
                                          ("sampler_tiny8_all", 8, ["length", "ss", "inpainting"]),
                                          ("sampler_tiny8_nocond", 8, [])])
def test_native_loop_matches_golden_reference_run(golden_dir, name, c, kinds):
    """The golden was produced by the unmodified reference with the numpy Philox stream; the in-kernel
    generator agrees with it to ~1e-6, so the fp32 engine lands on the same sample."""
    cfg, model, sd = make_native(tiny_cfg(c), "fp32")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, kinds) if kinds else {}
    sample, nfe = _run_native(cfg, model, ctx, cond, seed=2024)
    assert nfe == int(g["nfe"])
    _check_masks(sample, cond)
    assert rel_err(sample, torch.from_numpy(g["sample"])) < 1e-3


def test_graph_replay_equals_eager_and_generic_path():
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length"])
    a, _ = _run_native(cfg, model, ctx, cond, seed=7, use_graph=True)
    b, _ = _run_native(cfg, model, ctx, cond, seed=7, use_graph=False)
    assert torch.equal(a, b)

    class Wrapped(torch.nn.Module):  # not a UNetModel -> forces the generic update_fn loop
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, labels, context):
            return self.m(x, labels, context)

    c, nfe = _run_native(cfg, Wrapped(model), ctx, cond, seed=7)
    assert nfe == 8
    _check_masks(c, cond)
    # same update rules, different Philox stream ids -> compare against the oracle distributionally only
    assert torch.isfinite(c).all() and abs(c[:, :4].std().item() / a[:, :4].std().item() - 1) < 0.5


def test_truncated_run_and_sharding_independence():
    """Noise is keyed by the global sample index: sampling a batch of 4 in two shards of 2 gives the same
    noise; the states differ only through the batch-mean step size, i.e. each shard equals an
    oracle run of that shard."""
    cfg, model, sd = make_native(tiny_cfg(5, num_scales=10), "fp32")
    _, _, ctx = synthetic_inputs(cfg, 4, 8)
    cond = synthetic_condition(cfg, 4, ["length"])
    from text2protein_b200.score_sde_pytorch import sampling
    for r in range(2):
        sl = slice(2 * r, 2 * r + 2)
        cond_r = {"length": cond["length"][sl]}
        s, nfe = _run_native(cfg, model, ctx[sl], cond_r, seed=11, num_iters=3, sample_offset=2 * r)
        assert nfe == 6
        per = 5 * 32 * 32

        def noise_fn(stream, like, r=r):
            full = sampling.philox_normal((4, 5, 32, 32), 11, stream, "cuda").cpu()
            return full[2 * r: 2 * r + 2]

        sde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
        ref, _ = sampler_ref.pc_sampler_ref(sde, lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c),
                                            (2, 5, 32, 32), cfg.sampling.snr, n_steps=1, eps=1e-5, condition=cond_r,
                                            context=ctx[sl], noise_fn=noise_fn, num_iters=3)
        assert rel_err(s, ref) < 1e-4


def test_probability_flow_and_two_corrector_steps():
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    cfg.sampling.n_steps_each = 2
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    sample, nfe = _run_native(cfg, model, ctx, {}, seed=5)
    ref, ref_nfe = _run_oracle(cfg, sd, ctx, {}, seed=5)
    assert nfe == ref_nfe == 12
    assert rel_err(sample, ref) < 1e-4
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    fn = sampling.get_pc_sampler(sde, (2, 5, 32, 32), sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                 snr=0.17, n_steps=1, probability_flow=True, eps=1e-5, device="cuda", seed=5)
    s, _ = fn(model, {}, ctx.cuda())
    rsde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    r, _ = sampler_ref.pc_sampler_ref(
        rsde, lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), (2, 5, 32, 32), 0.17, n_steps=1,
        probability_flow=True, eps=1e-5, context=ctx,
        noise_fn=lambda stream, like: sampling.philox_normal(tuple(like.shape), 5, stream, "cuda").cpu())
    assert rel_err(s.cpu(), r) < 1e-4
