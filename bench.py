#!/usr/bin/env python
"""Headline benchmark: 6D maps/sec of the full predictor-corrector sampling loop (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE config 2, ``cond_length.yml`` -- N=128 6D maps, C=5, batch 64 per GPU,
length-conditioned, text context L=256 x 4096, VESDE num_scales=2000, snr 0.17, random-init re-randomised
weights, synthetic inputs.  A *step* is one PC iteration over the batch (Langevin corrector + reverse-diffusion
predictor = 2 score-network forwards + 2 fused step kernels).  ``value`` = maps/s of a full 2000-iteration run
extrapolated from the K timed iterations: n_gpus * B / (ms_per_step * num_scales).  Every rank runs the same
per-GPU batch (weak scaling); chains are independent, so there is no collective inside the loop.

The JSON line also carries: ``e2e`` (same metric through the public ``pc_sampler`` call with host buffers),
``roofline`` (dominant kernel, CUDA-event timed), ``cpu_baseline`` (the oracle port on the host cores),
``clocks`` (nvidia-smi during the timed region), ``gpu_launches``.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "6D maps/sec (N=128, full PC loop)"
# algorithmic FLOPs of one map at cond_length.yml = 2 * num_scales * 137.9 GFLOP (K/V projections hoisted)
FLOPS_PER_MAP = 551.7e12
UNIT = "maps/s"
WORKLOAD = "cond_length.yml N=128 C=5 B=64/GPU L=256 num_scales=2000 VESDE PC(langevin+reverse_diffusion)"
CTX_LEN = 256
BATCH = 64


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def _inputs(cfg, batch, seed=1234, sample_offset=0, want_ctx=True):
    g = torch.Generator().manual_seed(seed + sample_offset)
    N = cfg.data.max_res_num
    ctx = torch.randn(batch, CTX_LEN, cfg.model.context_dim, generator=g) * 0.02 if want_ctx else None
    lengths = torch.randint(40, N + 1, (batch,), generator=g)
    ar = torch.arange(N)
    lmask = (ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])
    return ctx, {"length": lmask}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def _oracle_setup(cfg, batch, seed_weights=42):
    """Oracle port of the reference path on CPU with the same re-randomised weights recipe and inputs."""
    import json as _json
    from oracle import sampler_ref, unet_ref
    tree = _json.load(open(os.path.join(ROOT, "tests", "golden", "param_tree_cond_length.json")))
    sd = unet_ref.state_dict_from_tree(tree, cfg, seed_weights)
    ctx, cond = _inputs(cfg, batch)
    sde = sampler_ref.VESDERef(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    shape = (batch, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    model = lambda x, lab, cx: unet_ref.unet_forward(sd, cfg, x, lab, cx)  # noqa: E731
    return sde, model, shape, ctx, cond


def _oracle_iterations(cfg, batch, iters, warm=1):
    """Seconds for `iters` PC iterations of the oracle port; weights / inputs are built and one warm-up iteration
    is run BEFORE the clock starts."""
    from oracle import sampler_ref
    sde, model, shape, ctx, cond = _oracle_setup(cfg, batch)

    def go(k):
        t0 = time.perf_counter()
        sampler_ref.pc_sampler_ref(sde, model, shape, cfg.sampling.snr, n_steps=cfg.sampling.n_steps_each, eps=1e-5,
                                   condition=cond, context=ctx, noise_fn=sampler_ref.philox_noise_fn(2024),
                                   num_iters=k)
        return time.perf_counter() - t0

    if warm:
        go(warm)
    return go(iters)


def run_reference(args, cfg):
    """--impl reference: the reference's CPU path (oracle port; the reference itself is Python and cannot travel
    to the GPU box) on all host cores, each step a bounded sample (1 map) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    batch = 1
    from oracle import sampler_ref
    sde, model, shape, ctx, cond = _oracle_setup(cfg, batch)
    noise = sampler_ref.philox_noise_fn(2024)

    def one_iter_run(k):
        t0 = time.perf_counter()
        sampler_ref.pc_sampler_ref(sde, model, shape, cfg.sampling.snr, n_steps=1, eps=1e-5, condition=cond,
                                   context=ctx, noise_fn=noise, num_iters=k)
        return time.perf_counter() - t0

    if args.warmup > 0:
        one_iter_run(args.warmup)
    dt = one_iter_run(args.steps)
    ms = dt / args.steps * 1e3
    value = batch / (ms * 1e-3 * cfg.model.num_scales)
    sample = f"B={batch} map(s), {args.steps} PC iterations of {cfg.model.num_scales}, fp32, torch CPU"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def _gemm_kernel_name(ksize, pixels_per_sample, n):
    """Which tcgen05 kernel the plan in csrc/gemm_tc.cu picks for a bf16 GEMM of this shape (make_plan)."""
    if n < 128:
        return "conv_gemm_tc_kernel"      # pixel-major
    if ksize == 3 and pixels_per_sample == 128 * 128:
        return "conv_gemm_tcH_kernel"     # halo variant: 3x3 on 128-pixel-wide images
    return "conv_gemm_tcT_kernel"         # channel-major


def _traffic_lookup(roof):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` captures (profiles/r01_traffic.json, keyed by the kernel string bench reports)."""
    key = roof["kernel"].split(" ", 1)[1]  # keyed by the shape part (+ " +gn" for launches that normalise their output)
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            v = json.load(open(p)).get(key)
            if v is not None:
                return v
    return None


def _step_kernel_roofline(dev, B, shape, L, mask_u8, hbm_gbs):
    """Predictor / corrector half-step kernels timed alone, configured as t2p_pc_run launches them (conditioned
    positions hold x_init from the start of the run and are not touched).  Algorithmic bytes per element
    (DESIGN.md 4.3), given the mask: a quad with a free position reads x (4), score (4), mask (1) and writes x (4)
    [+ x_mean (4) for the predictor]; a fully conditioned quad costs its mask byte; the corrector also reads the
    whole score once for the norms (4; the update's second read of it hits L2)."""
    from text2protein_b200 import _lib
    E = shape[1] * shape[2] * shape[3]
    n = B * E
    sets = 8  # 8 x 63 MB > L2
    xs = [torch.randn(shape, device=dev) for _ in range(sets)]
    sc = [torch.randn(shape, device=dev) for _ in range(sets)]
    xi = [torch.randn(shape, device=dev) for _ in range(sets)]
    xm = [torch.empty(shape, device=dev) for _ in range(sets)]
    G = torch.full((B,), 0.3, device=dev)
    ws = torch.empty(max(1, L.t2p_corrector_workspace_bytes(B, E) // 8), dtype=torch.float64, device=dev)
    free_quads = mask_u8.reshape(-1, 4).any(-1).float().mean().item()  # fraction of quads with a free position
    all_free = torch.ones_like(mask_u8)

    def timed(fn, predictor, mask):
        args = []
        for i in range(sets):
            a = _lib.StepArgs()
            a.x, a.score = xs[i].data_ptr(), sc[i].data_ptr()
            a.score_dtype, a.score_nhwc = 0, 0
            a.G = G.data_ptr()
            a.snr = 0.17
            a.mask, a.x_init = mask.data_ptr(), xi[i].data_ptr()
            a.conditioned_in_place = 1
            a.x_mean_out = xm[i].data_ptr() if predictor else None
            a.seed, a.stream_id, a.sample_offset = 2024, 5, 0
            a.B, a.C, a.HW = B, shape[1], shape[2] * shape[3]
            a.workspace = ws.data_ptr()
            args.append(a)
        for a in args:  # warm-up, eager
            _lib.check(fn(C.byref(a), _lib.current_stream()))
        torch.cuda.synchronize()
        # the launches are captured into a CUDA graph so that the timed region holds kernels only (as in the
        # sampling loop, which replays a graph), not Python / ctypes launch overhead
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        with torch.cuda.graph(graph, stream=cap):
            for a in args:
                _lib.check(fn(C.byref(a), _lib.current_stream()))
        graph.replay()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * sets)

    out = {}
    for name, fn, free_b, cond_b in (("predictor_kernel", L.t2p_predictor_step, 17, 1),
                                     ("corrector_kernel", L.t2p_corrector_step, 13, 5)):
        ms = timed(fn, name == "predictor_kernel", mask_u8)
        nbytes = n * (free_quads * free_b + (1.0 - free_quads) * cond_b)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": hbm_gbs,
                     "frac": gbs / hbm_gbs, "bound": "hbm", "free_quad_fraction": free_quads,
                     "note": "graph-replayed launches back to back; inputs rotate over 8 sets (> L2); bytes count "
                             "what the bench's length mask leaves to update; all_free = the same kernel with every "
                             "position free (17 / 13 B per element)"}
        try:  # the same kernel with nothing conditioned: the number comparable across rounds
            ms_free = timed(fn, name == "predictor_kernel", all_free)
            out[name]["all_free"] = {"ms": ms_free, "algorithmic_bytes": n * free_b,
                                     "achieved_gbs": n * free_b / (ms_free * 1e-3) / 1e9,
                                     "frac": n * free_b / (ms_free * 1e-3) / 1e9 / hbm_gbs}
        except Exception as e:  # never lose the bench line to the extra measurement
            out[name]["all_free"] = {"error": str(e)[:200]}
    return out


def _gn_apply_roofline(dev, B, cfg, L, hbm_gbs):
    """GroupNorm-apply + SiLU (the largest HBM-bound kernel of the forward, 18 % of its time) on the biggest tensor of
    the network, [B, N, N, nf] bf16: 2 B read + 2 B written per element."""
    from text2protein_b200 import _lib
    N, nf = cfg.data.max_res_num, cfg.model.nf
    sets = 3  # 3 x (268 + 268 MB) > L2
    xs = [torch.randn(B, N, N, nf, device=dev).bfloat16() for _ in range(sets)]
    ys = [torch.empty_like(x) for x in xs]
    scale = 1 + 0.1 * torch.randn(B, nf, device=dev)
    shift = 0.1 * torch.randn(B, nf, device=dev)
    def launch(i):
        _lib.check(L.t2p_groupnorm_apply(_lib.ptr(xs[i]), nf, None, 0, B, N, N, 1, _lib.ptr(scale), _lib.ptr(shift), 1, 0,
                                         _lib.ptr(ys[i]), None, _lib.current_stream()))
    for i in range(sets):
        launch(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=dev)):
        for i in range(sets):
            launch(i)
    graph.replay()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * sets)
    nbytes = xs[0].numel() * 4
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"gn_apply_rows_kernel": {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": hbm_gbs,
                                     "frac": gbs / hbm_gbs, "bound": "hbm",
                                     "note": f"[{B},{N},{N},{nf}] bf16, graph-replayed back to back over 3 buffer sets (> L2)"}}


# ------------------------------------------------------------------------------------------------ extra runs
def _time_loop(model, cfg, B, K, W, ctx_len, kinds, dev, sample_offset=0):
    """ms per PC iteration (CUDA events on the launching stream) of t2p_pc_run at per-GPU batch B: device-resident
    synthetic inputs, W warm-up iterations (eager + graph capture), K timed graph replays."""
    from text2protein_b200 import _lib
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    L = _lib.lib()
    C_, N = cfg.data.num_channels, cfg.data.max_res_num
    shape = (B, C_, N, N)
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    ctx = torch.randn(B, ctx_len, cfg.model.context_dim, device=dev) * 0.02
    cond = {"length": _inputs(cfg, B, sample_offset=sample_offset, want_ctx=False)[1]["length"].to(dev)} \
        if kinds == ["length"] else {}
    x = sampling.philox_normal(shape, 2024, 0, dev, scale=float(sde.sigma_max), sample_offset=sample_offset)
    x, cmask = sampling.apply_condition(x, cond)
    x = x.contiguous()
    x_init, x_mean = x.clone(), torch.empty_like(x)
    mask_u8 = cmask.contiguous().view(torch.uint8)
    model.set_context(ctx)
    labels, G = sampling.ve_tables(sde, 1e-5, max(K, W, 1))

    def run(k):
        a = _lib.RunArgs()
        a.x, a.x_mean, a.mask, a.x_init = x.data_ptr(), x_mean.data_ptr(), mask_u8.data_ptr(), x_init.data_ptr()
        a.label_table, a.g_table = labels.data_ptr(), G.data_ptr()
        a.num_iters, a.n_steps, a.snr, a.probability_flow = k, 1, float(cfg.sampling.snr), 0
        a.seed, a.sample_offset, a.B, a.use_graph = 2024, sample_offset, B, 1
        _lib.check(L.t2p_pc_run(model.native_handle, C.byref(a), _lib.current_stream()))

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        run(max(W, 2))
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        run(K)
        e1.record(side)
        side.synchronize()
    assert torch.isfinite(x).all()
    return e0.elapsed_time(e1) / K


def _extra_runs(args, dev, world, rank, cfg2_model, cfg2):
    """What BASELINE.json's configs 3-5 and the scaling question ask for beyond the headline line:
    (i) strong scaling of a FIXED global batch of cond_length.yml (64 and 256 maps over the run's N GPUs),
    (ii) test_config_large.yml with a global batch of 256 sharded over the N GPUs, (iii) the no_cond.yml sweep of
    8 ... 1024 chains.  Every number is ms per PC iteration, max over ranks, and whole-job maps/s; the per-N
    efficiency is for the reader (the driver) to form from the runs at N = 1, 2, 4, 8."""
    import torch.distributed as dist
    from text2protein_b200 import load_config
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel
    from text2protein_b200.synthetic import rerandomize_device_

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def row(name, cfg, model, total, ctx_len, kinds, K, note):
        b = total // world
        if b < 1:
            return {"workload": name, "global_batch": total, "skipped": "fewer chains than GPUs"}
        ms = reduce_max(_time_loop(model, cfg, b, K, 2, ctx_len, kinds, dev, sample_offset=rank * b))
        ns = cfg.model.num_scales
        return {"workload": name, "global_batch": b * world, "batch_per_gpu": b, "ms_per_iteration": ms,
                "maps_per_s": b * world / (ms * 1e-3 * ns), "iterations_timed": K, "limiter": note(b)}

    def gemm_or_launch(b):
        return ("tensor-core GEMMs (power-capped)" if b >= 32 else
                "launch-/latency-bound low-resolution tail: ~400 kernels per forward at a few microseconds each "
                "regardless of the batch")

    rows = []
    for total in (64, 256):
        rows.append(row("strong scaling: cond_length.yml, fixed global batch", cfg2, cfg2_model, total, CTX_LEN,
                        ["length"], 6, gemm_or_launch))

    def fresh(name):
        cfg = load_config(name, device=f"cuda:{dev.index}")
        cfg.model.compute_dtype = "bf16"
        with torch.device(dev):
            m = UNetModel(cfg)
        rerandomize_device_(m.named_parameters(), 42)
        m.sync_weights()
        return cfg, m

    if not args.no_cfg5:
        cfg5, m5 = fresh("no_cond")
        for total in (8, 64, 256, 1024):
            if total // world > 1024:
                continue
            rows.append(row("no_cond.yml sweep (C=8, unconditional, context L=256)", cfg5, m5, total, CTX_LEN, [], 4,
                            gemm_or_launch))
        del m5
        torch.cuda.empty_cache()
    if not args.no_cfg4:
        cfg4, m4 = fresh("test_config_large")
        rows.append(row("test_config_large.yml N=256 nf=256 L=512, global batch 256 sharded", cfg4, m4, 256, 512, [], 3,
                        lambda b: "tensor-core GEMMs; 22 attention pairs at T<=1024, d_head=128"))
        del m4
        torch.cuda.empty_cache()
    return rows


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args, cfg):
    import torch.distributed as dist
    from text2protein_b200.synthetic import rerandomize_
    from text2protein_b200 import _lib
    from text2protein_b200.distributed import gather_samples
    from text2protein_b200.score_sde_pytorch import sampling, sde_lib
    from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg.device = f"cuda:{local}"
    cfg.model.compute_dtype = "bf16"
    B = BATCH
    torch.manual_seed(cfg.seed)
    model = UNetModel(cfg).to(dev)
    rerandomize_(model.named_parameters(), 42)
    model.sync_weights()

    ctx_h, cond_h = _inputs(cfg, B, sample_offset=rank * B)
    ctx_pin = ctx_h.pin_memory()
    len_pin = cond_h["length"].pin_memory()
    sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
    shape = (B, cfg.data.num_channels, cfg.data.max_res_num, cfg.data.max_res_num)
    K, W = args.steps, args.warmup
    L = _lib.lib()

    # ---------------- device-resident loop (value): persistent buffers, graph-replayed iterations
    ctx_d = ctx_pin.to(dev, non_blocking=True)
    cond_d = {"length": len_pin.to(dev, non_blocking=True)}
    x = sampling.philox_normal(shape, 2024, 0, dev, scale=float(sde.sigma_max), sample_offset=rank * B)
    x, cmask = sampling.apply_condition(x, cond_d)
    x = x.contiguous()
    x_init = x.clone()
    x_mean = torch.empty_like(x)
    mask_u8 = cmask.contiguous().view(torch.uint8)
    model.set_context(ctx_d)
    labels, G = sampling.ve_tables(sde, 1e-5, max(K, W, 1))

    def run(k):
        a = _lib.RunArgs()
        a.x, a.x_mean, a.mask, a.x_init = x.data_ptr(), x_mean.data_ptr(), mask_u8.data_ptr(), x_init.data_ptr()
        a.label_table, a.g_table = labels.data_ptr(), G.data_ptr()
        a.num_iters, a.n_steps, a.snr, a.probability_flow = k, 1, float(cfg.sampling.snr), 0
        a.seed, a.sample_offset, a.B, a.use_graph = 2024, rank * B, B, 1
        _lib.check(L.t2p_pc_run(model.native_handle, C.byref(a), _lib.current_stream()))

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        run(max(W, 1))  # warm-up: first iteration eager, graph captured, the rest replayed
        side.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        run(K)
        e1.record(side)
        side.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clk = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / K
    num_scales = cfg.model.num_scales
    value = world * B / (ms_step * 1e-3 * num_scales)
    launches_fwd = int(L.t2p_unet_launches_per_forward(model.native_handle))
    gpu_launches = K * (2 * launches_fwd + 3)

    # ---------------- end to end through the public API (e2e): host buffers in, samples out, every run
    sampler = sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                      snr=cfg.sampling.snr, n_steps=1, eps=1e-5, device=cfg.device, seed=2024,
                                      num_iters=K, sample_offset=rank * B)
    out_pin = torch.empty(shape, dtype=torch.float32).pin_memory()
    ctx_in = torch.empty_like(ctx_pin, device=dev)                   # a serving loop's persistent input buffers
    len_in = torch.empty_like(len_pin, device=dev)

    def e2e_once():
        ctx_in.copy_(ctx_pin, non_blocking=True)                     # H2D: text context (268 MB fp32)
        len_in.copy_(len_pin, non_blocking=True)                     # H2D: length mask
        c, cd = ctx_in, {"length": len_in}
        s, _ = sampler(model, cd, c)
        if world > 1:                                                # final NCCL all-gather of the samples
            gather_samples(s, world * B)
        out_pin.copy_(s, non_blocking=True)                          # D2H: this rank's maps
        torch.cuda.synchronize()

    e2e_once()  # warm-up (captures the graph for these buffers)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_once()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_step = t.item() * 1e3 / K
    e2e_value = world * B / (e2e_ms_step * 1e-3 * num_scales)
    h2d = (ctx_pin.numel() * 4 + len_pin.numel()) / K
    d2h = out_pin.numel() * 4 / K

    # ---------------- roofline of the dominant kernel: CUDA events around every implicit-GEMM launch of eager
    # forward passes (the same launches the graph replays), aggregated per kernel shape
    roof = None
    fwd_flops = None
    if rank == 0:
        peak_burst, peak_sust, hbm, src = _peaks()
        xs = x.clone()
        lab = torch.full((B,), 7, dtype=torch.int64, device=dev)
        model(xs, lab, ctx_d)  # warm
        _lib.check(L.t2p_unet_set_profile(model.native_handle, 1))
        reps = 3
        for _ in range(reps):
            model(xs, lab, ctx_d)
        torch.cuda.synchronize()
        recs = (_lib.GemmRecord * 4096)()
        n = L.t2p_unet_profile_read(model.native_handle, recs, 4096)
        _lib.check(L.t2p_unet_set_profile(model.native_handle, 0))
        groups = {}
        total_flops = 0.0
        for r in recs[:n]:
            fl = 2.0 * r.M * r.N * r.K
            total_flops += fl
            key = (r.tensor_core, r.ksize, r.M, r.N, r.K)
            g = groups.setdefault(key, [0, 0.0, fl])
            g[0] += 1
            g[1] += r.ms
        fwd_flops = total_flops / reps
        key, (cnt, ms, fl) = max(groups.items(), key=lambda kv: kv[1][1])
        avg_ms = ms / cnt
        achieved = fl / (avg_ms * 1e-3) / 1e12
        tc_ms = sum(v[1] for k, v in groups.items() if k[0]) / reps
        all_ms = sum(v[1] for v in groups.values()) / reps
        # tensor_core == 2: the launch also applied the consumer's GroupNorm + SiLU to its output (one kernel where the
        # reference runs conv, GroupNorm and SiLU): its time is reported against the GEMM's FLOPs alone
        shapes = [{"k": k[1], "M": k[2], "N": k[3], "K": k[4], "n": v[0] // reps, "ms_each": v[1] / v[0],
                   "tflops": v[2] / (v[1] / v[0] * 1e-3) / 1e12, "epilogue": "groupnorm+silu" if k[0] == 2 else "plain"}
                  for k, v in sorted(groups.items(), key=lambda kv: -kv[1][1])[:16]]
        per_gpu = value / world
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                "frac": achieved / peak_burst, "frac_of_sustained": achieved / peak_sust, "traffic": None,
                # whole loop: maps/s per GPU x algorithmic FLOPs per map over the same peak
                "loop_achieved": per_gpu * FLOPS_PER_MAP / 1e12, "loop_frac": per_gpu * FLOPS_PER_MAP / 1e12 / peak_burst,
                "loop_frac_of_sustained": per_gpu * FLOPS_PER_MAP / 1e12 / peak_sust,
                "kernel": f"{_gemm_kernel_name(key[1], key[2] // B, key[3])} k={key[1]} M={key[2]} N={key[3]} K={key[4]}"
                          + (" +gn" if key[0] == 2 else ""),
                "kernel_note": ("this launch also applies the consumer's GroupNorm + SiLU to its own output "
                                "(epilogue GroupNorm, DESIGN 4.2c); FLOPs counted are the GEMM's alone") if key[0] == 2 else None,
                "launches_per_forward": cnt // reps, "avg_launch_ms": avg_ms,
                "flops_per_launch": fl,
                "peak_source": f"{src}: burst {peak_burst} (per-launch event timing), sustained {peak_sust}",
                "share_of_gemm_time": ms / reps / all_ms,
                "gemm_ms_per_forward": all_ms, "tc_gemm_ms_per_forward": tc_ms,
                "gemm_flops_per_forward": fwd_flops,
                "forward_tflops_incl_everything": (2 * fwd_flops) / (ms_step * 1e-3) / 1e12,
                "gemm_shapes_by_time": shapes}

    # ---------------- fused PC half-step kernels against the HBM roofline: CUDA events around eager launches,
    # rotating over buffer sets larger than the 126 MB L2 so that every launch streams from HBM
    steps = None
    if rank == 0:
        steps = _step_kernel_roofline(dev, B, shape, L, mask_u8, hbm)
        steps.update(_gn_apply_roofline(dev, B, cfg, L, hbm))
        if roof is not None:
            roof["traffic"] = _traffic_lookup(roof)

    # ---------------- CPU baseline: the oracle port on this box's host cores, bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        iters = 6
        dt = _oracle_iterations(cfg, 1, iters)
        cpu_ms = dt / iters * 1e3
        cpu = {"value": 1 / (cpu_ms * 1e-3 * num_scales), "unit": UNIT, "cores": torch.get_num_threads(),
               "kind": "port", "ms_per_step": cpu_ms,
               "sample": f"B=1 map, {iters} PC iterations of {num_scales} after 1 warm-up iteration, set-up excluded, "
                         "fp32 torch CPU"}
    if world > 1 and not args.no_cpu_baseline:
        # the other ranks sleep on the rendezvous store (a blocking socket wait) while rank 0 uses the host cores:
        # an NCCL barrier here would spin one core per rank
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("t2p_cpu_baseline_done", "1")
        else:
            store.wait(["t2p_cpu_baseline_done"])

    extra = None
    if not args.no_extra:
        try:
            extra = _extra_runs(args, dev, world, rank, model, cfg)
        except Exception as e:  # never lose the headline line to the extra measurements
            extra = [{"error": repr(e)[:300]}]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "l2": "working set (GBs of activations per "
                           "forward) exceeds the 126 MB L2; no explicit flush", "extrapolated_from_iterations": K,
                           "num_scales": num_scales, "parallelism": f"dp{world} (independent chains, no collective in the loop)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_step},
                "gpu_launches": gpu_launches, "launches_per_forward": launches_fwd,
                "score_net_forward_ms": (ms_step - sum(v["ms"] for k, v in (steps or {}).items()
                                                       if k in ("predictor_kernel", "corrector_kernel"))) / 2,
                "roofline": roof, "step_kernels": steps, "cpu_baseline": cpu, "clocks": clk,
                "extra_runs": extra, "library": os.path.basename(_lib.LIB_PATH),
                "env_t2p": {k: v for k, v in os.environ.items() if k.startswith("T2P_")},
                "workspace_gb": L.t2p_unet_workspace_bytes(model.native_handle) / 1e9}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: keep a private handle on the real stdout for it and point fd 1 at
    stderr, so that anything a library prints there (NCCL's version banner under NCCL_DEBUG, for one) cannot get
    in front of the line."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling / cfg4 / cfg5 runs")
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--lib", default="default",
                    help="default = libt2p.so; knobs = libt2p_knobs.so (-DT2P_TIMING_KNOBS build: reads the T2P_* A/B "
                         "environment variables); or the file name of another build under text2protein_b200/")
    args = ap.parse_args()
    if args.lib != "default":
        from text2protein_b200 import _lib
        _lib.use_library("libt2p_knobs.so" if args.lib == "knobs" else args.lib)
    from text2protein_b200 import load_config
    cfg = load_config("cond_length", device="cpu")
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for "
                             "the CPU baseline)")
        run_native(args, cfg)


if __name__ == "__main__":
    main()
