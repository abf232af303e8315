"""One eager score-network forward at the bench workload, bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists and single-kernel captures (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200.synthetic import rerandomize_  # noqa: E402
from text2protein_b200 import load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402

if os.environ.get("T2P_LIB"):  # e.g. libt2p_knobs.so (the -DT2P_TIMING_KNOBS build, which reads the T2P_* A/B variables)
    _lib.use_library(os.environ["T2P_LIB"])
B = int(os.environ.get("T2P_B", "64"))
cfg = load_config("cond_length", device="cuda")
cfg.model.compute_dtype = "bf16"
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, 5, 128, 128, generator=g) * 10).cuda()
ctx = (torch.randn(B, 256, 4096, generator=g) * 0.02).cuda()
lab = torch.full((B,), 7, dtype=torch.int64, device="cuda")
out = model(x, lab, ctx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = model(x, lab, ctx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(out.abs().mean()))
