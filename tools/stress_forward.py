"""Stress: the same score-network forward over and over (eager), every output compared bit for bit with the first.
    python tools/stress_forward.py [seconds] [B] [config]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
name = sys.argv[3] if len(sys.argv) > 3 else "cond_length"
cfg = load_config(name, device="cuda")
cfg.model.compute_dtype = "bf16"
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
g = torch.Generator().manual_seed(0)
N, C = cfg.data.max_res_num, cfg.data.num_channels
x = (torch.randn(B, C, N, N, generator=g) * 10).cuda()
ctx = (torch.randn(B, 256, cfg.model.context_dim, generator=g) * 0.02).cuda()
lab = torch.full((B,), 7, dtype=torch.int64, device="cuda")
ref = model(x, lab, ctx).clone()
torch.cuda.synchronize()
t0, n, bad = time.time(), 0, 0
while time.time() - t0 < secs:
    outs = [model(x, lab, ctx) for _ in range(1)]
    if not torch.equal(outs[0], ref):
        bad += 1
    n += 1
torch.cuda.synchronize()
print(f"{n} forwards at B = {B} ({name}), {bad} differed from the first")
assert bad == 0
