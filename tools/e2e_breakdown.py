"""Where the fixed cost of one public pc_sampler call goes (bench workload, K iterations)."""
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

K = int(os.environ.get("K", "10"))
cfg = load_config("cond_length", device="cuda")
cfg.model.compute_dtype = "bf16"
B = 64
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
ctx_h, cond_h = bench._inputs(cfg, B)
ctx_pin, len_pin = ctx_h.pin_memory(), cond_h["length"].pin_memory()
sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
shape = (B, 5, 128, 128)
fn = sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, snr=0.17,
                             n_steps=1, eps=1e-5, device="cuda", seed=2024, num_iters=K)
out_pin = torch.empty(shape).pin_memory()


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    t0 = T()
    c = ctx_pin.to("cuda", non_blocking=True)
    cd = {"length": len_pin.to("cuda", non_blocking=True)}
    t1 = T()
    model.sync_weights()
    t2 = T()
    model.set_context(c)
    t3 = T()
    s, _ = fn(model, cd, c)
    t4 = T()
    out_pin.copy_(s, non_blocking=True)
    t5 = T()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.1f} ms | sync_weights {1e3*(t2-t1):.1f} | set_context {1e3*(t3-t2):.1f} | "
          f"sampler(K={K}) {1e3*(t4-t3):.1f} = {1e3*(t4-t3)/K:.2f}/iter | d2h {1e3*(t5-t4):.1f}")
