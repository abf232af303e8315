"""One bf16 forward of a BASELINE configuration (debugging aid).  python tools/debug_model.py <config> <B> <L> [knobs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib, load_config  # noqa: E402

if len(sys.argv) > 4:
    _lib.use_library("libt2p_knobs.so")
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_device_  # noqa: E402

name, B, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = load_config(name, device="cuda")
cfg.model.compute_dtype = "bf16"
with torch.device("cuda"):
    m = UNetModel(cfg)
rerandomize_device_(m.named_parameters(), 42)
g = torch.Generator(device="cuda").manual_seed(3)
C_, N = cfg.data.num_channels, cfg.data.max_res_num
x = torch.randn(B, C_, N, N, device="cuda", generator=g) * 5
labels = torch.randint(0, cfg.model.num_scales, (B,), device="cuda", generator=g)
ctx = torch.randn(B, L, cfg.model.context_dim, device="cuda", generator=g) * 0.02
out = m(x, labels, ctx)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), "finite", bool(torch.isfinite(out).all()))
