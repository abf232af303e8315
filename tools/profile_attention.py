"""Self-attention of one SpatialTransformer at the bench workload (64 samples x 8 heads x 256 tokens x d = 32) through
t2p_attention(use_tensor_cores = 2), for ncu captures of attention_tc_kernel (see profiles/README.md).
    python tools/profile_attention.py [heads d T Tk]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from text2protein_b200 import _lib  # noqa: E402

heads, d, T, Tk = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (8, 32, 256, 256)
B = 64
inner = heads * d
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B, T, 3 * inner, device="cuda", generator=g).bfloat16()
kv = torch.randn(B, Tk, 2 * inner, device="cuda", generator=g).bfloat16()
out = torch.empty(B, T, inner, dtype=torch.bfloat16, device="cuda")
L = _lib.lib()
for _ in range(3):
    _lib.check(L.t2p_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(kv.data_ptr()), C.c_void_p(kv.data_ptr() + inner * 2),
                               _lib.ptr(out), B, heads, T, Tk, d, 3 * inner, 2 * inner, 2 * inner, inner, d ** -0.5,
                               _lib.BF16, 2, _lib.current_stream()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    _lib.check(L.t2p_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(kv.data_ptr()), C.c_void_p(kv.data_ptr() + inner * 2),
                               _lib.ptr(out), B, heads, T, Tk, d, 3 * inner, 2 * inner, 2 * inner, inner, d ** -0.5,
                               _lib.BF16, 2, _lib.current_stream()))
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
flops = 4.0 * B * heads * T * Tk * d
print(f"attention_tc heads={heads} d={d} T={T} Tk={Tk}: {us:.1f} us per launch, {flops / us / 1e6:.1f} TFLOP/s, finite {bool(torch.isfinite(out.float()).all())}")
