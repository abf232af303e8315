"""Per-launch GEMM times of one eager forward (engine profile mode) with the epilogue GroupNorm on and off, side by side.
    python tools/gemm_ab.py [B] [library file, e.g. libt2p_knobs.so]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib, load_config  # noqa: E402
from text2protein_b200.score_sde_pytorch.models.ncsnpp import UNetModel  # noqa: E402
from text2protein_b200.synthetic import rerandomize_  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
if len(sys.argv) > 2:
    _lib.use_library(sys.argv[2])
cfg = load_config("cond_length", device="cuda")
cfg.model.compute_dtype = "bf16"
model = UNetModel(cfg).cuda()
rerandomize_(model.named_parameters(), 42)
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, 5, 128, 128, generator=g) * 10).cuda()
ctx = (torch.randn(B, 256, 4096, generator=g) * 0.02).cuda()
lab = torch.full((B,), 7, dtype=torch.int64, device="cuda")
L = _lib.lib()


def profile(on):
    model.set_epilogue_groupnorm(on)
    model(x, lab, ctx)
    torch.cuda.synchronize()
    _lib.check(L.t2p_unet_set_profile(model.native_handle, 1))
    reps = 3
    for _ in range(reps):
        model(x, lab, ctx)
    recs = (_lib.GemmRecord * 8192)()
    n = L.t2p_unet_profile_read(model.native_handle, recs, 8192)
    _lib.check(L.t2p_unet_set_profile(model.native_handle, 0))
    per = n // reps
    out = []
    for i in range(per):
        r = recs[i]
        ms = sum(recs[i + k * per].ms for k in range(reps)) / reps
        out.append(((r.H, r.W, r.N, r.K, r.ksize), ms))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(x, lab, ctx)
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / 5


off, t_off = profile(False)
on, t_on = profile(True)
print(f"eager forward: off {t_off:.3f} ms, on {t_on:.3f} ms; GEMM sum off {sum(m for _, m in off):.3f} on {sum(m for _, m in on):.3f}")
assert [k for k, _ in off] == [k for k, _ in on]
agg = {}
for (k, a), (_, b) in zip(off, on):
    e = agg.setdefault(k, [0, 0.0, 0.0, 0])
    e[0] += 1
    e[1] += a
    e[2] += b
    if abs(b - a) > 0.15 * a:
        e[3] += 1
print("H W N K ks : launches  off_ms  on_ms  (launches that moved > 15 %)")
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(*k, ":", e[0], f"{e[1]:.3f} {e[2]:.3f}", e[3])
