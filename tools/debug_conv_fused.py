"""Stand-alone launch of the fused-GroupNorm halo convolution for one shape (debugging aid: run under
compute-sanitizer).  python tools/debug_conv_fused.py B H c0 c1 xc0 cout [library file]"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402

B, H, c0, c1, xc0, cout = [int(v) for v in sys.argv[1:7]]
if len(sys.argv) > 7:
    _lib.use_library(sys.argv[7])
W = 128
g = torch.Generator(device="cuda").manual_seed(11)
bf = lambda t: t.bfloat16().float()  # noqa: E731
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().bfloat16()  # noqa: E731
a0 = bf(torch.randn(B, c0, H, W, device="cuda", generator=g))
a1 = bf(torch.randn(B, c1, H, W, device="cuda", generator=g)) if c1 else None
x0 = bf(torch.randn(B, xc0, H, W, device="cuda", generator=g)) if xc0 else None
ctot = c0 + c1
scale = 1.0 + 0.3 * torch.randn(B, ctot, device="cuda", generator=g)
shift = 0.5 * torch.randn(B, ctot, device="cuda", generator=g)
w1 = bf(torch.randn(cout, ctot, 3, 3, device="cuda", generator=g) / math.sqrt(9 * ctot))
cols = [w1.permute(0, 2, 3, 1).reshape(cout, -1)]
x = a0 if a1 is None else torch.cat([a0, a1], 1)
ref = F.conv2d(F.silu(x * scale[:, :, None, None] + shift[:, :, None, None]), w1, None, padding=1)
if xc0:
    w2 = bf(torch.randn(cout, xc0, 1, 1, device="cuda", generator=g) / math.sqrt(xc0))
    ref = ref + F.conv2d(x0, w2)
    cols.append(w2.reshape(cout, -1))
wp = torch.cat(cols, 1).bfloat16().contiguous()
A0, A1, X0 = nhwc(a0), (nhwc(a1) if c1 else None), (nhwc(x0) if xc0 else None)
out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
a = _lib.ConvArgs()
a.a0, a.c0 = A0.data_ptr(), c0
if c1:
    a.a1, a.c1 = A1.data_ptr(), c1
a.B, a.H, a.W, a.ksize = B, H, W, 3
a.w, a.N, a.alpha = wp.data_ptr(), cout, 1.0
if xc0:
    a.x0, a.xc0 = X0.data_ptr(), xc0
a.out, a.out_dtype, a.in_dtype = out.data_ptr(), _lib.BF16, _lib.BF16
L = _lib.lib()
print("fuses", L.t2p_conv2d_fuses_groupnorm(C.byref(a)))
a.gn_scale, a.gn_shift = scale.data_ptr(), shift.data_ptr()
_lib.check(L.t2p_conv2d(C.byref(a), _lib.current_stream()))
torch.cuda.synchronize()
got = out.float().permute(0, 3, 1, 2)
print("rel err", ((got - ref).abs().max() / ref.abs().max()).item())
