"""A few eager launches of the fused PC half-step kernels at the bench workload (B=64, C=5, N=128), for ncu."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402

B, Cc, N = 64, 5, 128
shape = (B, Cc, N, N)
dev = "cuda"
L = _lib.lib()
g = torch.Generator().manual_seed(0)
lengths = torch.randint(40, N + 1, (B,), generator=g)
ar = torch.arange(N)
lm = (ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])
mask = torch.ones(shape, dtype=torch.bool) * lm[:, None]
mask[:, -1] = False
mask_u8 = mask.to(dev).contiguous().view(torch.uint8)
G = torch.full((B,), 0.3, device=dev)
ws = torch.empty(max(1, L.t2p_corrector_workspace_bytes(B, Cc * N * N) // 8), dtype=torch.float64, device=dev)
for it in range(4):
    x, sc, xi, xm = (torch.randn(shape, device=dev) for _ in range(4))
    for fn, pred in ((L.t2p_corrector_step, False), (L.t2p_predictor_step, True)):
        a = _lib.StepArgs()
        a.x, a.score, a.score_dtype, a.score_nhwc = x.data_ptr(), sc.data_ptr(), 0, 0
        a.G, a.snr = G.data_ptr(), 0.17
        a.mask, a.x_init = mask_u8.data_ptr(), xi.data_ptr()
        a.conditioned_in_place = 1
        a.x_mean_out = xm.data_ptr() if pred else None
        a.seed, a.stream_id, a.sample_offset = 2024, 5, 0
        a.B, a.C, a.HW = B, Cc, N * N
        a.workspace = ws.data_ptr()
        _lib.check(fn(C.byref(a), _lib.current_stream()))
torch.cuda.synchronize()
print("ok")
