"""Fixed cost of one public pc_sampler call (host buffers in, maps out): wall time at two iteration counts -> intercept
and slope.    python tools/e2e_fixed_cost.py [sync]      (sync: the weight checksum is awaited before the run is queued)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from text2protein_b200 import load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch import sampling, sde_lib  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_  # noqa: E402

if "sync" in sys.argv[1:]:
    _orig = UNetModel.sync_weights
    UNetModel.sync_weights = lambda self, force=False, check_data=False: _orig(self, force, bool(check_data))
cfg = load_config("cond_length", device="cuda")
cfg.model.compute_dtype = "bf16"
B = 64
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
ctx_h, cond_h = bench._inputs(cfg, B)
ctx_pin, len_pin = ctx_h.pin_memory(), cond_h["length"].pin_memory()
sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
shape = (B, 5, 128, 128)
out_pin = torch.empty(shape).pin_memory()
ctx_in, len_in = torch.empty_like(ctx_pin, device="cuda"), torch.empty_like(len_pin, device="cuda")
res = {}
for K in (10, 30):
    fn = sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, snr=0.17,
                                 n_steps=1, eps=1e-5, device="cuda", seed=2024, num_iters=K)

    def once():
        ctx_in.copy_(ctx_pin, non_blocking=True)
        len_in.copy_(len_pin, non_blocking=True)
        s, _ = fn(model, {"length": len_in}, ctx_in)
        out_pin.copy_(s, non_blocking=True)
        torch.cuda.synchronize()

    once()
    once()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        once()
        ts.append(1e3 * (time.perf_counter() - t0))
    res[K] = min(ts)
    print(f"K={K}: {['%.2f' % t for t in ts]} ms")
slope = (res[30] - res[10]) / 20
print(f"per iteration {slope:.3f} ms, fixed cost per call {res[10] - 10 * slope:.2f} ms")
