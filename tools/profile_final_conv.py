"""A few launches of the fused final layer at the bench shape (B=64, 128x128, 128 -> 5), for ncu."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402

B, H, W, cin, nout = 64, 128, 128, 128, 5
g = torch.Generator(device="cuda").manual_seed(8)
X = torch.randn(B, H, W, cin, device="cuda", generator=g).bfloat16()
scale = 1 + 0.3 * torch.randn(B, cin, device="cuda", generator=g)
shift = 0.5 * torch.randn(B, cin, device="cuda", generator=g)
wp = (torch.randn(nout, 9 * cin, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16()
bias = torch.randn(nout, device="cuda", generator=g)
out = torch.empty(B, nout, H, W, device="cuda")
for _ in range(4):
    _lib.check(_lib.lib().t2p_final_conv(_lib.ptr(X), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(wp), _lib.ptr(bias),
                                         _lib.ptr(out), B, H, W, cin, nout, _lib.current_stream()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    _lib.check(_lib.lib().t2p_final_conv(_lib.ptr(X), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(wp), _lib.ptr(bias),
                                         _lib.ptr(out), B, H, W, cin, nout, _lib.current_stream()))
e1.record()
torch.cuda.synchronize()
print(f"ok {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
