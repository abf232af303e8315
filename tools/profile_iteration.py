"""One eager predictor-corrector iteration (2 score-network forwards + the two fused half-steps) at the bench
workload, bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch import sampling, sde_lib  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_  # noqa: E402

B, N = int(os.environ.get("T2P_B", "64")), 128
cfg = load_config("cond_length", device="cuda")
cfg.model.compute_dtype = "bf16"
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
g = torch.Generator().manual_seed(0)
ctx = (torch.randn(B, 256, 4096, generator=g) * 0.02).cuda()
lengths = torch.randint(40, N + 1, (B,), generator=g)
ar = torch.arange(N)
cond = {"length": ((ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])).cuda()}
sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)


def run(k):
    fn = sampling.get_pc_sampler(sde, (B, 5, N, N), sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                 snr=cfg.sampling.snr, n_steps=1, eps=1e-5, device="cuda", seed=7, num_iters=k,
                                 use_graph=False)
    out, nfe = fn(model, cond, ctx)
    torch.cuda.synchronize()
    return out


run(1)
torch.cuda.cudart().cudaProfilerStart()
out = run(1)
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(out.abs().mean()))
