"""Round-2 parity cases: VPSDE through the sampler and the step kernels, float time conditioning, the
symmetrisation option (default off; north_star d), reference-exact masking with several corrector steps, the CUDA
graph cache across runs of different length, EMA weight hand-over, and the opt-in global-batch step size of a sharded
run (needs two GPUs)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

from oracle import sampler_ref, unet_ref
from tests.cfgs import synthetic_condition, synthetic_inputs, tiny_cfg
from tests.gpu_util import make_native, rel_err
from text2protein_b200 import _lib

pytestmark = pytest.mark.gpu


def _st():
    return _lib.current_stream()


def _noise(seed, stream, shape):
    n = torch.empty(shape, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().t2p_philox_normal(C.c_uint64(seed), stream, 0, n.numel(), C.c_float(1.0), _lib.ptr(n), _st()))
    return n


def _to_dev(cond):
    return {k: ({a: b.cuda() for a, b in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in cond.items()}


def _sampler(cfg, sde, B, **kw):
    from text2protein_b200.score_sde_pytorch import sampling
    shape = (B, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    kw.setdefault("eps", 1e-5)
    kw.setdefault("n_steps", cfg.sampling.n_steps_each)
    return sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                   snr=cfg.sampling.snr, device="cuda", **kw)


# ------------------------------------------------------------------------------------------------ VPSDE (a9)
@pytest.mark.parametrize("pf", [False, True])
def test_vp_step_kernels_match_float64_math(pf):
    """t2p_predictor_step / t2p_corrector_step with the VP arguments (sqrt_alpha: f = sqrt(alpha) x - x,
    sde_lib.py:149-157; alpha-scaled Langevin step, sampling.py:184-186,195) against the reference's expressions
    evaluated in float64 on the device, with the kernel's own normals."""
    B, Cc, N, seed = 3, 5, 32, 41
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(B, Cc, N, N, generator=g) * 2).cuda()
    score = torch.randn(B, Cc, N, N, generator=g, dtype=torch.float64).cuda()
    ref_sde = sampler_ref.VPSDERef(0.1, 20.0, 50)
    ts = torch.tensor([49, 20, 0])
    beta, alpha = ref_sde.discrete_betas[ts].cuda(), ref_sde.alphas[ts].cuda()
    G, sa = torch.sqrt(beta).contiguous(), torch.sqrt(alpha).contiguous()
    L = _lib.lib()

    def args(stream):
        a = _lib.StepArgs()
        xb = x.clone()
        a.x, a.score, a.score_dtype = xb.data_ptr(), score.data_ptr(), _lib.F64
        a.seed, a.stream_id, a.B, a.C, a.HW = seed, stream, B, Cc, N * N
        return a, xb

    # predictor
    a, xp = args(5)
    xm = torch.empty_like(xp)
    a.G, a.sqrt_alpha, a.x_mean_out, a.probability_flow = G.data_ptr(), sa.data_ptr(), xm.data_ptr(), int(pf)
    _lib.check(L.t2p_predictor_step(C.byref(a), _st()))
    f = sa[:, None, None, None] * x - x
    rev_f = f.double() - (G[:, None, None, None] ** 2).double() * score * (0.5 if pf else 1.0)
    xm_ref = x.double() - rev_f
    xp_ref = xm_ref if pf else xm_ref + (G[:, None, None, None] * _noise(seed, 5, x.shape)).double()
    assert rel_err(xm, xm_ref.float()) < 2e-6 and rel_err(xp, xp_ref.float()) < 2e-6
    # corrector
    a, xc = args(6)
    ws = torch.empty(L.t2p_corrector_workspace_bytes(B, Cc * N * N) // 8, dtype=torch.float64, device="cuda")
    al = alpha.contiguous()
    a.alpha, a.snr, a.workspace = al.data_ptr(), 0.17, ws.data_ptr()
    _lib.check(L.t2p_corrector_step(C.byref(a), _st()))
    z = _noise(seed, 6, x.shape)
    gn = torch.norm(score.reshape(B, -1), dim=-1).mean()
    nn_ = torch.norm(z.reshape(B, -1), dim=-1).mean()
    step = (0.17 * nn_ / gn) ** 2 * 2 * al
    xc_ref = x + step[:, None, None, None] * score + torch.sqrt(step * 2)[:, None, None, None] * z
    assert rel_err(xc, xc_ref.float()) < 2e-6


def test_vpsde_sampler_matches_reference_golden(golden_dir):
    """VPSDE through get_pc_sampler (generic path: score_fn with FLOAT time conditioning -> native score network ->
    fused step kernels with the VP arguments) against the reference's own pc_sampler run and the oracle."""
    from text2protein_b200.score_sde_pytorch import sde_lib
    g = np.load(os.path.join(golden_dir, "sampler_vpsde_tiny5.npz"))
    N, K = int(g["N"]), int(g["K"])
    cfg, model, sd = make_native(tiny_cfg(5, num_scales=N), "fp32")
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length"])
    sde = sde_lib.VPSDE(cfg.model.beta_min, cfg.model.beta_max, N)
    fn = _sampler(cfg, sde, 2, eps=1e-3, seed=2024, num_iters=K)
    s, nfe = fn(model, _to_dev(cond), ctx.cuda())
    s = s.cpu()
    assert nfe == 2 * K and s.dtype == torch.float32
    assert torch.equal(s[:, -1], cond["length"].float())
    assert rel_err(s, torch.from_numpy(g["sample"])) < 1e-3
    ref, _ = sampler_ref.pc_sampler_ref(
        sampler_ref.VPSDERef(cfg.model.beta_min, cfg.model.beta_max, N),
        lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), (2, 5, 32, 32), cfg.sampling.snr, n_steps=1,
        eps=1e-3, condition=cond, context=ctx, num_iters=K, generic_streams=True,
        noise_fn=lambda stream, like: _noise(2024, stream, tuple(like.shape)).cpu())
    assert rel_err(s, ref) < 1e-4


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_float_time_conditioning_is_embedded_as_float(dtype, tol):
    """ncsnpp.py:221-223: the sinusoidal embedding takes the time conditioning as given (fractional for VP models),
    only the sigma lookup truncates it."""
    cfg, model, sd = make_native(tiny_cfg(5, num_scales=24), dtype)
    x, _, ctx = synthetic_inputs(cfg, 3, 8)
    t = torch.tensor([22.75, 7.5, 0.003])
    out = model(x.cuda(), t.cuda(), ctx.cuda())
    ref = unet_ref.unet_forward(sd, cfg, x, t, ctx)
    assert rel_err(out, ref) < tol
    trunc = model(x.cuda(), t.long().cuda(), ctx.cuda())
    assert rel_err(trunc, ref) > 1e-2  # the truncated label is a different (wrong) embedding: 4e-2 away


# ------------------------------------------------------------------------------------------------ symmetrisation
@pytest.mark.parametrize("kind", ["predictor", "corrector"])
def test_symmetrize_step_kernel_matches_specification(kind):
    B, Cc, N, seed = 2, 5, 32, 17
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(B, Cc, N, N, generator=g) * 4).cuda()
    h = torch.randn(B, Cc, N, N, generator=g).cuda()
    cond = synthetic_condition(tiny_cfg(5), B, ["length"])
    cm = (torch.ones(B, Cc, N, N, dtype=torch.bool) * cond["length"][:, None])
    cm[:, -1] = False
    cm[0, 0, 3, 5] = False  # one asymmetric hole: its transpose (5, 3) stays un-symmetrised too
    x_init = torch.randn(B, Cc, N, N, generator=g).cuda()
    G = (torch.rand(B, generator=g) + 0.2).cuda()
    L = _lib.lib()
    mask_u8 = cm.cuda().contiguous().view(torch.uint8)
    ws = torch.empty(L.t2p_corrector_workspace_bytes(B, Cc * N * N) // 8, dtype=torch.float64, device="cuda")

    def run(sym):
        a = _lib.StepArgs()
        out, xm = torch.empty_like(x), torch.empty_like(x)
        a.x, a.x_out, a.score, a.score_dtype = x.data_ptr(), out.data_ptr(), h.data_ptr(), _lib.F32
        a.G, a.snr, a.mask, a.x_init, a.x_mean_out = G.data_ptr(), 0.17, mask_u8.data_ptr(), x_init.data_ptr(), xm.data_ptr()
        a.seed, a.stream_id, a.B, a.C, a.HW, a.W = seed, 9, B, Cc, N * N, N
        a.workspace, a.symmetrize = ws.data_ptr(), int(sym)
        fn = L.t2p_predictor_step if kind == "predictor" else L.t2p_corrector_step
        _lib.check(fn(C.byref(a), _st()))
        torch.cuda.synchronize()
        return out.cpu(), xm.cpu()

    off, off_mean = run(False)
    on, on_mean = run(True)
    z = _noise(seed, 9, x.shape).double()
    if kind == "predictor":
        um = x.double() + (G[:, None, None, None] ** 2).double() * h.double()
        un = um + (G[:, None, None, None] * z.float()).double()
    else:
        gn = torch.norm(h.double().reshape(B, -1), dim=-1).mean()
        nn_ = torch.norm(z.float().reshape(B, -1), dim=-1).mean()
        step = ((0.17 * nn_ / gn) ** 2 * 2).float()
        um = x.double() + step.double() * h.double()
        un = um + (torch.sqrt(step * 2) * z.float()).double()
    for got_on, got_off, u in ((on, off, un), (on_mean, off_mean, um)):
        spec = torch.where(cm, sampler_ref.symmetrize_free(u.cpu(), cm), x_init.cpu().double()).float()
        assert rel_err(got_on, spec) < 2e-6
        assert rel_err(got_off, torch.where(cm, u.cpu(), x_init.cpu().double()).float()) < 2e-6
        both = cm[:, :2] & cm[:, :2].transpose(2, 3)
        sym = got_on[:, :2]
        assert torch.equal(sym[both], sym.transpose(2, 3)[both])          # EXACTLY symmetric where both are free
        assert torch.equal(got_on[:, 2:], got_off[:, 2:])                  # other channels: bit-identical to OFF
        assert torch.equal(got_on[~cm], x_init.cpu()[~cm])
        assert torch.equal(got_on[0, 0, 5, 3], got_off[0, 0, 5, 3])        # partner of the hole: left alone
    with pytest.raises(_lib.NativeError):                                  # in place + symmetrize is refused
        a = _lib.StepArgs()
        a.x, a.score, a.G, a.B, a.C, a.HW, a.W, a.symmetrize = x.data_ptr(), h.data_ptr(), G.data_ptr(), B, Cc, N * N, N, 1
        _lib.check(L.t2p_predictor_step(C.byref(a), _st()))


@pytest.mark.parametrize("kinds,n_steps", [(["length"], 1), ([], 2)])
def test_symmetrize_sampler_on_matches_oracle_and_off_is_the_reference(golden_dir, kinds, n_steps):
    from text2protein_b200.score_sde_pytorch import sde_lib
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    cfg.sampling.n_steps_each = n_steps
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, kinds) if kinds else {}
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    on, _ = _sampler(cfg, sde, 2, seed=2024, symmetrize=True)(model, _to_dev(cond), ctx.cuda())
    off, _ = _sampler(cfg, sde, 2, seed=2024)(model, _to_dev(cond), ctx.cuda())
    off2, _ = _sampler(cfg, sde, 2, seed=2024, symmetrize=False)(model, _to_dev(cond), ctx.cuda())
    on, off = on.cpu(), off.cpu()
    assert torch.equal(off, off2.cpu())
    if kinds == ["length"] and n_steps == 1:  # OFF is the path the reference golden pins
        g = np.load(os.path.join(golden_dir, "sampler_tiny5_length.npz"))
        assert rel_err(off, torch.from_numpy(g["sample"])) < 1e-3
    assert torch.equal(on[:, :2], on[:, :2].transpose(2, 3))  # symmetric masks: symmetric everywhere
    assert not torch.equal(off[:, :2], off[:, :2].transpose(2, 3))
    if "length" in cond:
        assert torch.equal(on[:, -1], cond["length"].float())
    ref, _ = sampler_ref.pc_sampler_ref(
        sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales),
        lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), (2, 5, 32, 32), cfg.sampling.snr, n_steps=n_steps,
        eps=1e-5, condition=cond, context=ctx, symmetrize=True,
        noise_fn=lambda stream, like: _noise(2024, stream, tuple(like.shape)).cpu())
    assert rel_err(on, ref) < 1e-4


def test_symmetrize_generic_path():
    """Same option through the generic update_fn loop (a wrapped model defeats the fast path)."""
    from text2protein_b200.score_sde_pytorch import sde_lib

    class Wrapped(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, labels, context):
            return self.m(x, labels, context)

    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length"])
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    s, _ = _sampler(cfg, sde, 2, seed=5, symmetrize=True)(Wrapped(model), _to_dev(cond), ctx.cuda())
    s = s.cpu()
    assert torch.equal(s[:, :2], s[:, :2].transpose(2, 3)) and torch.equal(s[:, -1], cond["length"].float())
    ref, _ = sampler_ref.pc_sampler_ref(
        sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales),
        lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), (2, 5, 32, 32), cfg.sampling.snr, n_steps=1,
        eps=1e-5, condition=cond, context=ctx, symmetrize=True, generic_streams=True,
        noise_fn=lambda stream, like: _noise(5, stream, tuple(like.shape)).cpu())
    assert rel_err(s, ref) < 1e-4


# ------------------------------------------------------------------------------------------------ loop details
def test_two_corrector_steps_with_a_condition_mask_like_the_reference():
    """sampling.py:282-283 applies the condition after the whole corrector update: with n_steps_each = 2 the first
    inner step runs unmasked and its drifted conditioned positions feed the second score evaluation."""
    from text2protein_b200.score_sde_pytorch import sde_lib
    cfg, model, sd = make_native(tiny_cfg(8), "fp32")
    cfg.sampling.n_steps_each = 2
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length", "ss", "inpainting"])
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    for denoise in (True, False):
        s, nfe = _sampler(cfg, sde, 2, seed=31, denoise=denoise)(model, _to_dev(cond), ctx.cuda())
        ref, ref_nfe = sampler_ref.pc_sampler_ref(
            sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales),
            lambda a, b, c: unet_ref.unet_forward(sd, cfg, a, b, c), (2, 8, 32, 32), cfg.sampling.snr, n_steps=2,
            eps=1e-5, condition=cond, context=ctx, denoise=denoise,
            noise_fn=lambda stream, like: _noise(31, stream, tuple(like.shape)).cpu())
        assert nfe == ref_nfe == 12
        assert rel_err(s, ref) < 1e-4
        keep = ~cond["inpainting"]["mask_inpaint"][:, None].expand_as(ref)
        assert torch.equal(s.cpu()[keep], cond["inpainting"]["coords_6d"][keep])


def test_graph_cache_survives_runs_of_different_length_and_batch():
    """A cached iteration graph holds raw pointers into the loop's tables and the network's arena: a longer run
    (tables re-allocated) or a bigger forward in between (arena re-allocated) must not replay a stale graph."""
    from text2protein_b200.score_sde_pytorch import sde_lib
    cfg, model, sd = make_native(tiny_cfg(5, num_scales=40), "fp32")
    _, _, ctx = synthetic_inputs(cfg, 2, 8)
    cond = synthetic_condition(cfg, 2, ["length"])
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    dev_cond, dev_ctx = _to_dev(cond), ctx.cuda()

    def run(K, graph):
        s, _ = _sampler(cfg, sde, 2, seed=3, num_iters=K, use_graph=graph)(model, dev_cond, dev_ctx)
        return s.cpu()

    a2 = run(2, True)
    a6 = run(6, True)      # same buffers (caching allocator), same seed: a cache hit with a different K
    assert torch.equal(a6, run(6, False)) and torch.equal(a2, run(2, False))
    a40 = run(40, True)    # more iterations than any earlier run
    x8, l8, c8 = synthetic_inputs(cfg, 8, 8)
    model(x8.cuda(), l8.cuda(), c8.cuda())  # grows the activation arena between two same-key runs
    assert torch.equal(run(40, True), a40) and torch.equal(a40, run(40, False))
    assert torch.equal(run(6, True), a6)


def test_ema_copy_to_and_data_writes_reach_the_engine():
    """store -> copy_to -> sample -> restore (the reference's evaluation idiom, models/ema.py:51-83)."""
    from text2protein_b200.score_sde_pytorch import sde_lib
    from text2protein_b200.score_sde_pytorch.models.ema import ExponentialMovingAverage
    cfg, model, sd = make_native(tiny_cfg(5), "fp32")
    x, labels, ctx = synthetic_inputs(cfg, 2, 8)
    x, labels, ctx = x.cuda(), labels.cuda(), ctx.cuda()
    ema = ExponentialMovingAverage(model.parameters(), decay=0.5)
    base = model(x, labels, ctx)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.05)
    ema.update(model.parameters())            # shadow = something between the two weight sets
    live = model(x, labels, ctx)
    assert not torch.equal(live, base)
    ema.store(model.parameters())
    ema.copy_to(model.parameters())
    sd_ema = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    out = model(x, labels, ctx)
    assert rel_err(out, unet_ref.unet_forward(sd_ema, cfg, x.cpu(), labels.cpu(), ctx.cpu())) < 1e-5
    ema.restore(model.parameters())
    assert torch.equal(model(x, labels, ctx), live)
    # a writer that bypasses the version counter (p.data.copy_, the reference's own idiom): caught by the sampler's
    # per-run checksum
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    s0, _ = _sampler(cfg, sde, 2, seed=1)(model, {}, ctx)
    for p, shadow in zip(model.parameters(), ema.shadow_params):
        p.data.copy_(shadow.data)
    s1, _ = _sampler(cfg, sde, 2, seed=1)(model, {}, ctx)
    assert not torch.equal(s0, s1)
    ref, _ = sampler_ref.pc_sampler_ref(
        sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales),
        lambda a, b, c: unet_ref.unet_forward(sd_ema, cfg, a, b, c), (2, 5, 32, 32), cfg.sampling.snr, n_steps=1,
        eps=1e-5, context=ctx.cpu(), noise_fn=lambda stream, like: _noise(1, stream, tuple(like.shape)).cpu())
    assert rel_err(s1, ref) < 1e-4


# ------------------------------------------------------------------------------------------------ sharded step size
def _sync_worker(rank, world, port, q):
    import torch.distributed as dist
    from text2protein_b200.distributed import StepSizeSync, gather_samples, shard_range
    from text2protein_b200.score_sde_pytorch import sde_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    total = 6
    cfg, model, sd = make_native(tiny_cfg(5, num_scales=10), "fp32")
    _, _, ctx = synthetic_inputs(cfg, total, 8)
    cond = synthetic_condition(cfg, total, ["length"])
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    a, b = shard_range(total, rank, world)
    sync = StepSizeSync(total)
    outs = []
    for rep in range(2):  # two runs through the same group: graph replay + fresh mailbox tags
        fn = _sampler(cfg, sde, b - a, seed=11, num_iters=6, sample_offset=a, sync_step_size=sync)
        s, _ = fn(model, {"length": cond["length"][a:b].cuda()}, ctx[a:b].cuda())
        outs.append(gather_samples(s, total))
    fn = _sampler(cfg, sde, b - a, seed=11, num_iters=6, sample_offset=a)  # default: per-shard step size
    s, _ = fn(model, {"length": cond["length"][a:b].cuda()}, ctx[a:b].cuda())
    local = gather_samples(s, total)
    if rank == 0:
        whole, _ = _sampler(cfg, sde, total, seed=11, num_iters=6)(model, {"length": cond["length"].cuda()}, ctx.cuda())
        q.put((outs[0].cpu().numpy(), outs[1].cpu().numpy(), local.cpu().numpy(), whole.cpu().numpy()))
    dist.barrier()
    sync.close()
    dist.destroy_process_group()


def test_sync_step_size_makes_sharded_run_equal_the_global_batch():
    """Two ranks x 3 samples with sync_step_size == one process sampling all 6 (the reference's batch-mean step
    size over the whole batch, sampling.py:193-195): fp32 <= 1e-5.  Without the option the shards differ."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        synced, synced2, local, whole = q.get(timeout=600)
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    synced, synced2, local, whole = map(torch.from_numpy, (synced, synced2, local, whole))
    assert rel_err(synced, whole) < 1e-5
    assert torch.equal(synced, synced2)
    assert rel_err(local, whole) > 1e-4  # per-shard step sizes: every shard is its own reference run
