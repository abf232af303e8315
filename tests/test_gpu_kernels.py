"""Per-kernel parity (through the C ABI) against torch fp32 / the oracle.  Needs a B200."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import philox_ref
from tests.gpu_util import nchw, nhwc, rel_err
from text2protein_b200 import _lib

pytestmark = pytest.mark.gpu

# references are plain fp32 math: keep cuDNN / cuBLAS off their TF32 paths
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _st():
    return _lib.current_stream()


def _conv(a0, w, ksize, bias=None, a1=None, rowbias=None, residual=None, res_up=0, alpha=1.0, in_dtype=torch.float32,
          out_dtype=None, stats=False):
    """a0/a1 NCHW fp32 on cuda, w [Cout, Cin, k, k] fp32.  Returns NCHW fp32 (and stats)."""
    out_dtype = out_dtype or in_dtype
    B, c0, H, W = a0.shape
    c1 = a1.shape[1] if a1 is not None else 0
    N = w.shape[0]
    wp = w.permute(0, 2, 3, 1).contiguous().reshape(N, -1).to(in_dtype).contiguous()
    A0 = nhwc(a0, in_dtype)
    A1 = nhwc(a1, in_dtype) if a1 is not None else None
    out = torch.empty(B, H, W, N, dtype=out_dtype, device="cuda")
    R = None
    if residual is not None:
        R = nhwc(residual, out_dtype)
    a = _lib.ConvArgs()
    a.a0, a.c0 = A0.data_ptr(), c0
    a.a1, a.c1 = (A1.data_ptr() if A1 is not None else None), c1
    a.B, a.H, a.W, a.ksize = B, H, W, ksize
    a.w, a.N = wp.data_ptr(), N
    a.bias = bias.data_ptr() if bias is not None else None
    if rowbias is not None:
        a.rowbias, a.rowbias_ld = rowbias.data_ptr(), rowbias.shape[1]
    a.residual = R.data_ptr() if R is not None else None
    a.res_up, a.alpha = res_up, alpha
    a.out, a.out_dtype, a.in_dtype = out.data_ptr(), _lib.torch_dtype_code(out_dtype), _lib.torch_dtype_code(in_dtype)
    ss = None
    if stats:
        tile = _lib.lib().t2p_conv2d_stat_tile(C.byref(a))
        assert tile > 0 and (H * W) % tile == 0, tile
        ss = torch.full((B, H * W // tile, N, 2), float("nan"), dtype=torch.float32, device="cuda")
        a.stat_part = ss.data_ptr()
    _lib.check(_lib.lib().t2p_conv2d(C.byref(a), _st()))
    torch.cuda.synchronize()
    return nchw(out.float()), ss


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("H,cin,cin1,cout,k", [(32, 64, 0, 128, 3), (16, 128, 64, 64, 3), (8, 128, 0, 128, 3),
                                               (4, 64, 0, 192, 3), (16, 64, 128, 64, 1), (64, 64, 0, 5, 3),
                                               (128, 64, 64, 128, 3), (128, 128, 0, 256, 3),
                                               # one pixel per sample: the per-sample bias changes inside a pixel pair
                                               (1, 128, 0, 256, 1)])
def test_conv2d_matches_torch(dtype, tol, H, cin, cin1, cout, k):
    g = torch.Generator(device="cuda").manual_seed(1)
    B = 3
    a0 = torch.randn(B, cin, H, H, device="cuda", generator=g)
    a1 = torch.randn(B, cin1, H, H, device="cuda", generator=g) if cin1 else None
    w = torch.randn(cout, cin + cin1, k, k, device="cuda", generator=g) / math.sqrt((cin + cin1) * k * k)
    bias = torch.randn(cout, device="cuda", generator=g)
    rowbias = torch.randn(B, cout + 7, device="cuda", generator=g)
    res = torch.randn(B, cout, H, H, device="cuda", generator=g)
    if dtype == torch.bfloat16:  # compare on identical (bf16-representable) operands
        a0, w, res = a0.bfloat16().float(), w.bfloat16().float(), res.bfloat16().float()
        a1 = a1.bfloat16().float() if a1 is not None else None
    x = a0 if a1 is None else torch.cat([a0, a1], 1)
    ref = (F.conv2d(x, w, bias, padding=k // 2) + rowbias[:, :cout, None, None] + res) * 0.70710678
    out, _ = _conv(a0, w, k, bias=bias, a1=a1, rowbias=rowbias, residual=res, alpha=0.70710678, in_dtype=dtype)
    assert rel_err(out, ref) < tol


@pytest.mark.parametrize("B,H,cin,cout,k", [(8, 4, 256, 256, 3), (64, 4, 512, 256, 3), (8, 16, 256, 256, 3), (3, 8, 256, 384, 3),
                                            (1, 16, 512, 256, 3), (2, 4, 2048, 256, 1), (5, 2, 128, 200, 3)])
def test_conv2d_split_k(B, H, cin, cout, k):
    """Launches of few tiles and long K (the 16 x 16 ... 4 x 4 levels) share the k-blocks of a tile out over several CTAs
    and the last arrival adds the fp32 partials in split order: against torch on the same bf16 operands, with the
    per-sample bias changing inside a pixel tile (4 x 4 images), a residual through the identity k-blocks, a ragged
    channel tile and fused statistics; two calls give the same bits (the sum does not depend on who arrives last)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    a0 = bf(torch.randn(B, cin, H, H, device="cuda", generator=g))
    w = bf(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
    bias = torch.randn(cout, device="cuda", generator=g)
    rowbias = torch.randn(B, cout + 3, device="cuda", generator=g)
    res = bf(torch.randn(B, cout, H, H, device="cuda", generator=g))
    ref = (F.conv2d(a0, w, bias, padding=k // 2) + rowbias[:, :cout, None, None] + res) * 0.70710678
    stats = (H * H) % 32 == 0 and cout % 128 == 0
    out, ss = _conv(a0, w, k, bias=bias, rowbias=rowbias, residual=res, alpha=0.70710678, in_dtype=torch.bfloat16, stats=stats)
    assert rel_err(out, ref) < 1.5e-2
    out2, ss2 = _conv(a0, w, k, bias=bias, rowbias=rowbias, residual=res, alpha=0.70710678, in_dtype=torch.bfloat16, stats=stats)
    assert torch.equal(out, out2)
    if stats:
        tot = ss.sum(dim=1)
        assert torch.allclose(tot[..., 0], out.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(tot[..., 1], (out * out).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
        assert torch.equal(ss, ss2)


@pytest.mark.parametrize("B,H,cin,cout,k", [(16, 32, 128, 256, 3), (8, 64, 64, 256, 3), (64, 16, 256, 256, 3), (256, 8, 64, 256, 1),
                                            (12, 32, 64, 512, 3), (1024, 4, 128, 256, 3)])
def test_conv2d_cta_pair(B, H, cin, cout, k):
    """Layers of 256 output channels and more at >= 96 tiles run on CTA pairs (tcgen05.mma.cta_group::2: M = 256 over two
    CTAs, each staging half of the 256-pixel tile): against torch on the same bf16 operands, with bias, per-sample bias,
    a residual through the identity k-blocks (the follower's identity block is shifted by 128 columns) and statistics."""
    g = torch.Generator(device="cuda").manual_seed(9)
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    a0 = bf(torch.randn(B, cin, H, H, device="cuda", generator=g))
    w = bf(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
    bias = torch.randn(cout, device="cuda", generator=g)
    rowbias = torch.randn(B, cout + 1, device="cuda", generator=g)
    res = bf(torch.randn(B, cout, H, H, device="cuda", generator=g))
    ref = (F.conv2d(a0, w, bias, padding=k // 2) + rowbias[:, :cout, None, None] + res) * 0.70710678
    stats = (H * H) % 128 == 0
    out, ss = _conv(a0, w, k, bias=bias, rowbias=rowbias, residual=res, alpha=0.70710678, in_dtype=torch.bfloat16, stats=stats)
    assert torch.isfinite(out).all()
    per = (out - ref).flatten(1).abs().amax(1) / ref.flatten(1).abs().amax(1)
    assert per.max() < 1.5e-2, per
    # every 128-channel tile on its own: a swapped pair would show here
    for c in range(0, cout, 128):
        assert rel_err(out[:, c:c + 128], ref[:, c:c + 128]) < 1.5e-2
    if stats:
        tot = ss.sum(dim=1)
        assert torch.allclose(tot[..., 0], out.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    out2, _ = _conv(a0, w, k, bias=bias, rowbias=rowbias, residual=res, alpha=0.70710678, in_dtype=torch.bfloat16, stats=stats)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("H,cin,cout", [(16, 64, 64), (16, 64, 128), (32, 128, 256), (64, 64, 128)])
def test_conv2d_upsampled_residual_and_fused_stats(H, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(2)
    B = 2
    a0 = torch.randn(B, cin, H, H, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16().float()
    res = torch.randn(B, cout, H // 2, H // 2, device="cuda", generator=g).bfloat16().float()
    ref = F.conv2d(a0, w, None, padding=1) + res.repeat_interleave(2, 2).repeat_interleave(2, 3)
    out, ss = _conv(a0, w, 3, residual=res, res_up=1, in_dtype=torch.bfloat16, stats=True)
    assert rel_err(out, ref) < 1.5e-2
    # statistics describe the tensor exactly as stored
    tot = ss.sum(dim=1)  # [B, N, 2] over the pixel tiles of each sample
    assert torch.allclose(tot[..., 0], out.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[..., 1], (out * out).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("B,H,c0,c1,xc0,cout,stats", [(2, 128, 128, 0, 0, 128, True), (3, 128, 128, 128, 0, 128, False),
                                                      (2, 128, 64, 0, 64, 128, True), (16, 6, 64, 64, 128, 256, False),
                                                      (5, 128, 128, 0, 128, 128, True), (1, 128, 512, 0, 512, 512, True),
                                                      (1, 128, 256, 0, 256, 512, False), (2, 128, 512, 256, 768, 256, True)])
def test_conv2d_fused_groupnorm_silu(B, H, c0, c1, xc0, cout, stats):
    """Conv3x3(silu(GroupNorm-affine(cat(a0, a1)))) [+ folded 1x1 skip over RAW x0] with the normalisation applied
    inside the halo kernel's operand path (packed-bf16 arithmetic) against torch on the same bf16 inputs, including
    image borders (zero padding applies AFTER the activation), ragged tile counts and fused output statistics."""
    g = torch.Generator(device="cuda").manual_seed(11)
    W = 128
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    a0 = bf(torch.randn(B, c0, H, W, device="cuda", generator=g) * 1.5 + 0.3)
    a1 = bf(torch.randn(B, c1, H, W, device="cuda", generator=g)) if c1 else None
    x0 = bf(torch.randn(B, xc0, H, W, device="cuda", generator=g)) if xc0 else None
    ctot = c0 + c1
    scale = 1.0 + 0.3 * torch.randn(B, ctot, device="cuda", generator=g)
    shift = 0.5 * torch.randn(B, ctot, device="cuda", generator=g)
    w1 = bf(torch.randn(cout, ctot, 3, 3, device="cuda", generator=g) / math.sqrt(9 * ctot))
    w2 = bf(torch.randn(cout, xc0, 1, 1, device="cuda", generator=g) / math.sqrt(xc0)) if xc0 else None
    bias = torch.randn(cout, device="cuda", generator=g)
    x = a0 if a1 is None else torch.cat([a0, a1], 1)
    hn = F.silu(x * scale[:, :, None, None] + shift[:, :, None, None])
    ref = F.conv2d(hn, w1, bias, padding=1)
    cols = [w1.permute(0, 2, 3, 1).reshape(cout, -1)]
    if xc0:
        ref = ref + F.conv2d(x0, w2)
        cols.append(w2.reshape(cout, -1))
    wp = torch.cat(cols, 1).bfloat16().contiguous()
    A0 = nhwc(a0, torch.bfloat16)
    A1 = nhwc(a1, torch.bfloat16) if a1 is not None else None
    X0 = nhwc(x0, torch.bfloat16) if x0 is not None else None
    out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    a = _lib.ConvArgs()
    a.a0, a.c0 = A0.data_ptr(), c0
    if A1 is not None:
        a.a1, a.c1 = A1.data_ptr(), c1
    a.B, a.H, a.W, a.ksize = B, H, W, 3
    a.w, a.N, a.bias, a.alpha = wp.data_ptr(), cout, bias.data_ptr(), 1.0
    if X0 is not None:
        a.x0, a.xc0 = X0.data_ptr(), xc0
    a.out, a.out_dtype, a.in_dtype = out.data_ptr(), _lib.BF16, _lib.BF16
    L = _lib.lib()
    assert L.t2p_conv2d_fuses_groupnorm(C.byref(a)) == 1
    a.gn_scale, a.gn_shift = scale.data_ptr(), shift.data_ptr()
    ss = None
    if stats:
        tile = L.t2p_conv2d_stat_tile(C.byref(a))
        assert tile > 0
        ss = torch.zeros(B * H * W // tile, cout, 2, device="cuda")
        a.stat_part = ss.data_ptr()
    _lib.check(L.t2p_conv2d(C.byref(a), _st()))
    torch.cuda.synchronize()
    got = nchw(out.float())
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 1.5e-2
    # border rows / columns exercise the zero padding of the ACTIVATION (silu(shift) != 0 would leak in otherwise)
    assert rel_err(got[:, :, 0], ref[:, :, 0]) < 1.5e-2 and rel_err(got[:, :, :, -1], ref[:, :, :, -1]) < 1.5e-2
    if stats:
        tot = ss.reshape(B, -1, cout, 2).sum(dim=1)
        assert torch.allclose(tot[..., 0], got.sum(dim=(2, 3)), rtol=1e-4, atol=1e-1)
    # a launch that cannot fuse says so and refuses the arguments
    b = _lib.ConvArgs()
    b.a0, b.c0, b.B, b.H, b.W, b.ksize, b.w, b.N = A0.data_ptr(), c0, B, H, W, 1, wp.data_ptr(), cout
    b.out, b.out_dtype, b.in_dtype = out.data_ptr(), _lib.BF16, _lib.BF16
    assert L.t2p_conv2d_fuses_groupnorm(C.byref(b)) == 0


@pytest.mark.parametrize("B,H,cin,cout,k", [(3, 128, 128, 128, 3), (20, 128, 64, 128, 3), (2, 64, 128, 256, 3),
                                            (5, 32, 256, 256, 3), (4, 16, 256, 256, 3), (2, 8, 128, 256, 3),
                                            (37, 64, 64, 128, 3), (3, 32, 128, 128, 1), (64, 16, 64, 512, 3),
                                            (6, 16, 128, 1024, 1), (48, 32, 128, 256, 3), (40, 16, 128, 1024, 1),
                                            # deferred form with four channel tiles and groups of 16 (the N = 256 configs)
                                            (2, 128, 64, 512, 3), (3, 128, 64, 256, 3)])
def test_conv2d_normalises_its_own_output(B, H, cin, cout, k):
    """silu(GroupNorm_1(Conv_0(h) + bias + temb)) (layers.py:314-318) from the convolution's own epilogue: per-sample
    statistics are exchanged between the CTAs of the launch.  Against torch on the same bf16 operands; B = 20 / 37 / 64
    put several waves of tiles through the persistent CTAs (samples that straddle two waves), and a second call on
    fresh buffers checks that the exchange words were handed back zeroed."""
    g = torch.Generator(device="cuda").manual_seed(21)
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    groups = min(cout // 4, 32)
    a0 = bf(torch.randn(B, cin, H, H, device="cuda", generator=g))
    w = bf(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
    bias = torch.randn(cout, device="cuda", generator=g)
    rowbias = torch.randn(B, cout + 5, device="cuda", generator=g)
    gamma = 1.0 + 0.3 * torch.randn(cout, device="cuda", generator=g)
    beta = 0.5 * torch.randn(cout, device="cuda", generator=g)
    # per-sample scale so that the statistics differ from sample to sample
    a0 = bf(a0 * (0.5 + torch.arange(B, device="cuda").float()[:, None, None, None] / B))
    conv = F.conv2d(a0, w, bias, padding=k // 2) + rowbias[:, :cout, None, None]
    ref = F.silu(F.group_norm(conv, groups, gamma, beta, eps=1e-6))
    wp = w.permute(0, 2, 3, 1).contiguous().reshape(cout, -1).bfloat16().contiguous()
    A0 = nhwc(a0, torch.bfloat16)
    L = _lib.lib()
    for rep in range(2):
        out = torch.full((B, H, H, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        a = _lib.ConvArgs()
        a.a0, a.c0 = A0.data_ptr(), cin
        a.B, a.H, a.W, a.ksize = B, H, H, k
        a.w, a.N, a.bias, a.alpha = wp.data_ptr(), cout, bias.data_ptr(), 1.0
        a.rowbias, a.rowbias_ld = rowbias.data_ptr(), rowbias.shape[1]
        a.out, a.out_dtype, a.in_dtype = out.data_ptr(), _lib.BF16, _lib.BF16
        a.gno_groups, a.gno_eps = groups, 1e-6
        if L.t2p_conv2d_normalises_output(C.byref(a)) != 1:
            # launches of few tiles split K over the idle SMs instead (and leave GroupNorm to the one-launch kernel)
            assert B * H * H * ((cout + 127) // 128) <= 74 * 256
            pytest.skip("this shape splits K")
        a.gno_gamma, a.gno_beta = gamma.data_ptr(), beta.data_ptr()
        _lib.check(L.t2p_conv2d(C.byref(a), _st()))
        torch.cuda.synchronize()
        got = nchw(out.float())
        assert torch.isfinite(got).all()
        assert rel_err(got, ref) < 1e-2
        # every sample on its own: a statistic taken from the wrong sample would show here
        per = (got - ref).flatten(1).abs().amax(1) / ref.flatten(1).abs().amax(1)
        assert per.max() < 1.5e-2, per
    # launches that cannot do it say so and refuse the arguments: odd group size, residual, too many tiles per sample
    b = _lib.ConvArgs()
    b.a0, b.c0, b.B, b.H, b.W, b.ksize, b.w, b.N = A0.data_ptr(), cin, B, H, H, k, wp.data_ptr(), cout
    b.out, b.out_dtype, b.in_dtype = out.data_ptr(), _lib.BF16, _lib.BF16
    b.gno_groups = 3
    assert L.t2p_conv2d_normalises_output(C.byref(b)) == 0
    b.gno_groups, b.residual = groups, out.data_ptr()
    assert L.t2p_conv2d_normalises_output(C.byref(b)) == 0
    b.gno_gamma, b.gno_beta = gamma.data_ptr(), beta.data_ptr()
    assert L.t2p_conv2d(C.byref(b), _st()) != 0


@pytest.mark.parametrize("H,cin,xc0,xc1,cout", [(16, 64, 128, 64, 128), (32, 128, 128, 0, 128), (8, 64, 64, 0, 256),
                                                (128, 64, 64, 64, 128)])
def test_conv2d_folded_skip_path(H, cin, xc0, xc1, cout):
    """Conv_1(h) + Conv_2(cat(x0, x1)) [+ residual] as one launch (skip path as centre-tap K columns)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    B = 3
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    h = bf(torch.randn(B, cin, H, H, device="cuda", generator=g))
    x0 = bf(torch.randn(B, xc0, H, H, device="cuda", generator=g))
    x1 = bf(torch.randn(B, xc1, H, H, device="cuda", generator=g)) if xc1 else None
    w1 = bf(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin))
    w2 = bf(torch.randn(cout, xc0 + xc1, 1, 1, device="cuda", generator=g) / math.sqrt(xc0 + xc1))
    bias = torch.randn(cout, device="cuda", generator=g)
    res = bf(torch.randn(B, cout, H, H, device="cuda", generator=g))
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = (F.conv2d(h, w1, bias, padding=1) + F.conv2d(x, w2) + res) * 0.70710678
    wp = torch.cat([w1.permute(0, 2, 3, 1).reshape(cout, -1), w2.reshape(cout, -1)], 1).bfloat16().contiguous()
    A, X0 = nhwc(h, torch.bfloat16), nhwc(x0, torch.bfloat16)
    X1 = nhwc(x1, torch.bfloat16) if x1 is not None else None
    R = nhwc(res, torch.bfloat16)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    a = _lib.ConvArgs()
    a.a0, a.c0, a.c1 = A.data_ptr(), cin, 0
    a.B, a.H, a.W, a.ksize = B, H, H, 3
    a.w, a.N, a.bias = wp.data_ptr(), cout, bias.data_ptr()
    a.x0, a.xc0 = X0.data_ptr(), xc0
    if X1 is not None:
        a.x1, a.xc1 = X1.data_ptr(), xc1
    a.residual, a.alpha = R.data_ptr(), 0.70710678
    a.out, a.out_dtype, a.in_dtype = out.data_ptr(), _lib.torch_dtype_code(torch.bfloat16), _lib.torch_dtype_code(torch.bfloat16)
    _lib.check(_lib.lib().t2p_conv2d(C.byref(a), _st()))
    torch.cuda.synchronize()
    assert rel_err(nchw(out.float()), ref) < 1.5e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,H,W,c0,c1,groups,silu", [(2, 16, 16, 256, 0, 32, 1), (3, 32, 32, 256, 256, 32, 1),
                                                     (2, 8, 8, 256, 0, 32, 0), (5, 4, 4, 256, 256, 32, 1),
                                                     (2, 16, 16, 64, 64, 32, 1), (2, 32, 32, 512, 512, 32, 0),
                                                     (1, 5, 7, 128, 0, 16, 1), (2, 16, 16, 1024, 0, 32, 1)])
def test_groupnorm_small_one_launch(dtype, tol, B, H, W, c0, c1, groups, silu):
    """Statistics + finalize + apply in one launch (what the engine runs at 32 x 32 and below) against torch:
    groups of 4 / 8 / 16 / 32 channels, two-source concat, ragged pixel counts, SiLU on and off."""
    g = torch.Generator(device="cuda").manual_seed(4)
    a0 = (torch.randn(B, c0, H, W, device="cuda", generator=g) * 2 + 0.5).to(dtype).float()
    a1 = (torch.randn(B, c1, H, W, device="cuda", generator=g) * 0.5 - 1.0).to(dtype).float() if c1 else None
    C_ = c0 + c1
    gamma = 1 + 0.1 * torch.randn(C_, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C_, device="cuda", generator=g)
    x = a0 if a1 is None else torch.cat([a0, a1], 1)
    ref = F.group_norm(x, groups, gamma, beta, eps=1e-6)
    if silu:
        ref = F.silu(ref)
    A0, A1 = nhwc(a0, dtype), (nhwc(a1, dtype) if a1 is not None else None)
    out = torch.empty(B, H, W, C_, dtype=dtype, device="cuda")
    _lib.check(_lib.lib().t2p_groupnorm_small(_lib.ptr(A0), c0, _lib.ptr(A1), c1, B, H, W, _lib.torch_dtype_code(dtype),
                                              groups, 1e-6, _lib.ptr(gamma), _lib.ptr(beta), silu, _lib.ptr(out), _st()))
    torch.cuda.synchronize()
    assert rel_err(nchw(out.float()), ref) < tol
    out2 = torch.empty_like(out)
    _lib.check(_lib.lib().t2p_groupnorm_small(_lib.ptr(A0), c0, _lib.ptr(A1), c1, B, H, W, _lib.torch_dtype_code(dtype),
                                              groups, 1e-6, _lib.ptr(gamma), _lib.ptr(beta), silu, _lib.ptr(out2), _st()))
    assert torch.equal(out, out2)  # deterministic
    # shapes outside its envelope are refused, not mis-computed
    rc = _lib.lib().t2p_groupnorm_small(_lib.ptr(A0), c0, _lib.ptr(A1), c1, B, 64, 64, _lib.torch_dtype_code(dtype),
                                        groups, 1e-6, _lib.ptr(gamma), _lib.ptr(beta), silu, _lib.ptr(out), _st())
    assert rc != 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("c0,c1,groups", [(64, 0, 16), (128, 64, 32), (256, 128, 32)])
def test_groupnorm_silu_resample_concat(dtype, tol, mode, c0, c1, groups):
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H = 2, 16
    a0 = (torch.randn(B, c0, H, H, device="cuda", generator=g) * 2 + 0.5).to(dtype).float()
    a1 = (torch.randn(B, c1, H, H, device="cuda", generator=g) - 1.0).to(dtype).float() if c1 else None
    C_ = c0 + c1
    gamma = 1 + 0.1 * torch.randn(C_, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C_, device="cuda", generator=g)
    x = a0 if a1 is None else torch.cat([a0, a1], 1)
    ref = F.silu(F.group_norm(x, groups, gamma, beta, eps=1e-6))
    raw_ref = None
    if mode == 1:
        ref = F.avg_pool2d(ref, 2)
        raw_ref = F.avg_pool2d(x, 2)
    elif mode == 2:
        ref = ref.repeat_interleave(2, 2).repeat_interleave(2, 3)
    OH = H // 2 if mode == 1 else (H * 2 if mode == 2 else H)
    A0, A1 = nhwc(a0, dtype), (nhwc(a1, dtype) if a1 is not None else None)
    out = torch.empty(B, OH, OH, C_, dtype=dtype, device="cuda")
    raw = torch.empty(B, OH, OH, C_, dtype=dtype, device="cuda") if mode == 1 else None
    _lib.check(_lib.lib().t2p_groupnorm(_lib.ptr(A0), c0, _lib.ptr(A1), c1, B, H, H, _lib.torch_dtype_code(dtype),
                                        groups, 1e-6, _lib.ptr(gamma), _lib.ptr(beta), 1, mode, _lib.ptr(out),
                                        _lib.ptr(raw), _st()))
    torch.cuda.synchronize()
    assert rel_err(nchw(out.float()), ref) < tol
    if raw is not None:
        assert rel_err(nchw(raw.float()), raw_ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("Cdim", [64, 256, 512, 1024])
def test_layernorm_and_geglu(dtype, tol, Cdim):
    g = torch.Generator(device="cuda").manual_seed(4)
    M = 77
    x = (torch.randn(M, Cdim, device="cuda", generator=g) * 1.5 + 0.3).to(dtype)
    gamma = 1 + 0.1 * torch.randn(Cdim, device="cuda", generator=g)
    beta = 0.1 * torch.randn(Cdim, device="cuda", generator=g)
    y = torch.empty_like(x)
    _lib.check(_lib.lib().t2p_layernorm(_lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta), M, Cdim, 1e-5,
                                        _lib.torch_dtype_code(dtype), _lib.ptr(y), _st()))
    assert rel_err(y.float(), F.layer_norm(x.float(), (Cdim,), gamma, beta)) < tol
    z = torch.randn(M, 2 * Cdim, device="cuda", generator=g).to(dtype)
    o = torch.empty(M, Cdim, dtype=dtype, device="cuda")
    _lib.check(_lib.lib().t2p_geglu(_lib.ptr(z), M, Cdim, _lib.torch_dtype_code(dtype), _lib.ptr(o), _st()))
    a, gate = z.float().chunk(2, dim=-1)
    assert rel_err(o.float(), a * F.gelu(gate)) < tol


# tc: 0 CUDA-core kernel, 1 mma.sync kernel, 2 tcgen05 / TMEM / TMA kernel (what the engine runs).  Shapes: the
# AttnBlockpp head (1 x 256, and 1 x 512 / 1 x 1024 of the N = 256 configs: value dimension split over CTAs), self
# attention (8 x 32 at T = 256; 8 x 64 / 8 x 128 at T = 1024: online-softmax rescale over 8 key blocks), cross
# attention to L = 77 / 256 / 512 text tokens (ragged last key block), mid blocks at T = 16 / 64 (ragged query tile)
@pytest.mark.parametrize("dtype,tc,tol", [(torch.float32, 0, 2e-5), (torch.bfloat16, 0, 2e-2), (torch.bfloat16, 1, 2e-2),
                                          (torch.bfloat16, 2, 2e-2)])
@pytest.mark.parametrize("heads,d,Tq,Tk", [(1, 256, 64, 64), (8, 32, 256, 256), (8, 32, 256, 77), (4, 16, 16, 8),
                                           (8, 64, 100, 300), (1, 512, 32, 32), (8, 128, 64, 512), (1, 256, 256, 256),
                                           (8, 32, 16, 16), (8, 64, 1024, 1024), (8, 128, 1024, 512),
                                           (1, 1024, 64, 64), (4, 32, 64, 8), (8, 32, 16, 256), (1, 128, 256, 256)])
def test_attention(dtype, tc, tol, heads, d, Tq, Tk):
    if tc == 1 and d % 16:
        pytest.skip("tensor-core kernel needs d % 16 == 0")
    if tc == 2 and d == 16:
        pytest.skip("d = 16 heads (tiny test network only) stay on the mma.sync kernel")
    g = torch.Generator(device="cuda").manual_seed(5)
    B = 2
    inner = heads * d
    # strided views, exactly like the fused QKV projection output
    qkv = torch.randn(B, Tq, 3 * inner, device="cuda", generator=g).to(dtype)
    kv = torch.randn(B, Tk, 2 * inner, device="cuda", generator=g).to(dtype)
    q = qkv[..., :inner]
    k, v = kv[..., :inner], kv[..., inner:]
    out = torch.empty(B, Tq, inner, dtype=dtype, device="cuda")
    scale = d ** -0.5
    es = qkv.element_size()
    _lib.check(_lib.lib().t2p_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(kv.data_ptr()),
                                        C.c_void_p(kv.data_ptr() + inner * es), _lib.ptr(out), B, heads, Tq, Tk, d,
                                        3 * inner, 2 * inner, 2 * inner, inner, scale, _lib.torch_dtype_code(dtype),
                                        tc, _st()))

    def split(t):
        return t.float().reshape(B, -1, heads, d).permute(0, 2, 1, 3)

    ref = torch.softmax(split(q) @ split(k).transpose(-1, -2) * scale, dim=-1) @ split(v)
    ref = ref.permute(0, 2, 1, 3).reshape(B, Tq, inner)
    assert rel_err(out.float(), ref) < tol


def test_philox_bits_exact_and_normals_close():
    seed, stream = 0x1234567ABCDEF, 7
    bits = torch.empty(4096 * 4, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().t2p_philox_bits(C.c_uint64(seed), stream, 1000, 4096, _lib.ptr(bits), _st()))
    got = bits.cpu().numpy().view(np.uint32).reshape(-1, 4)
    want = philox_ref.philox_bits(seed, stream, 1000, 4096)
    assert np.array_equal(got, want)  # integer path: bit-exact
    n = torch.empty(1 << 16, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().t2p_philox_normal(C.c_uint64(seed), stream, 4000, n.numel(), C.c_float(1.0), _lib.ptr(n),
                                            _st()))
    ref = philox_ref.philox_normal(seed, stream, 4000, n.numel())
    assert np.abs(n.cpu().numpy() - ref).max() < 2e-5


def _ref_steps(x, score64, G, snr, noise_c, noise_p, mask, x_init):
    """oracle restatement of one corrector + one predictor half-step (sampler_ref.pc_sampler_ref body)."""
    B = x.shape[0]
    grad = score64
    grad_norm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
    noise_norm = torch.norm(noise_c.reshape(B, -1), dim=-1).mean()
    step = (snr * noise_norm / grad_norm) ** 2 * 2 * torch.ones(B)
    xm = x + step[:, None, None, None] * grad
    xc = xm + torch.sqrt(step * 2)[:, None, None, None] * noise_c
    xc = torch.where(mask, xc, x_init).float()
    rev_f = torch.zeros_like(x) - G[:, None, None, None] ** 2 * score64
    xm2 = x - rev_f
    xp = xm2 + G[:, None, None, None] * noise_p
    return xc, torch.where(mask, xp, x_init).float(), torch.where(mask, xm2, x_init).float()


@pytest.mark.parametrize("B,Cc,N", [(2, 5, 32), (3, 8, 64), (64, 5, 16)])
def test_pc_steps_match_oracle_math(B, Cc, N):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, Cc, N, N, generator=g) * 10
    h = torch.randn(B, N, N, Cc, generator=g)                     # raw NHWC fp32 network output
    labels = torch.randint(0, 50, (B,), generator=g)
    sigmas = torch.tensor(np.exp(np.linspace(np.log(100.0), np.log(0.01), 50)))
    G = torch.rand(B, generator=g) + 0.1
    mask = torch.rand(B, Cc, N, N, generator=g) > 0.3
    x_init = torch.randn(B, Cc, N, N, generator=g)
    score64 = h.permute(0, 3, 1, 2).double() / sigmas[labels][:, None, None, None]
    seed = 99

    def noise(stream):
        n = torch.empty(B, Cc, N, N, dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib().t2p_philox_normal(C.c_uint64(seed), stream, 0, n.numel(), C.c_float(1.0), _lib.ptr(n),
                                                _st()))
        return n.cpu()

    xc_ref, xp_ref, xm_ref = _ref_steps(x, score64, G, 0.17, noise(11), noise(12), mask, x_init)
    dev = {k: v.cuda() for k, v in dict(h=h, labels=labels, sigmas=sigmas, G=G, x_init=x_init).items()}
    mask_u8 = mask.cuda().contiguous().view(torch.uint8)
    ws = torch.empty(_lib.lib().t2p_corrector_workspace_bytes(B, Cc * N * N) // 8, dtype=torch.float64, device="cuda")

    def args(xbuf, stream):
        a = _lib.StepArgs()
        a.x, a.score, a.score_dtype, a.score_nhwc = xbuf.data_ptr(), dev["h"].data_ptr(), _lib.F32, 1
        a.sigmas, a.labels, a.G = dev["sigmas"].data_ptr(), dev["labels"].data_ptr(), dev["G"].data_ptr()
        a.snr, a.mask, a.x_init = 0.17, mask_u8.data_ptr(), dev["x_init"].data_ptr()
        a.seed, a.stream_id, a.B, a.C, a.HW = seed, stream, B, Cc, N * N
        a.workspace = ws.data_ptr()
        return a

    xc = x.cuda().clone()
    _lib.check(_lib.lib().t2p_corrector_step(C.byref(args(xc, 11)), _st()))
    xp = x.cuda().clone()
    xm = torch.empty_like(xp)
    a = args(xp, 12)
    a.x_mean_out = xm.data_ptr()
    _lib.check(_lib.lib().t2p_predictor_step(C.byref(a), _st()))
    torch.cuda.synchronize()
    for got, ref in ((xc, xc_ref), (xp, xp_ref), (xm, xm_ref)):
        got = got.cpu()
        assert torch.equal(got[~mask], x_init[~mask])            # mask handling is bit-exact
        assert rel_err(got, ref) < 2e-6


def _length_mask(B, Cc, N, g, lo):
    """conditional_mask of a length condition (utils.py:62-81 shape): free inside the L x L corner, padding channel fixed."""
    lengths = torch.randint(lo, N + 1, (B,), generator=g)
    ar = torch.arange(N)
    lm = (ar[None, :, None] < lengths[:, None, None]) & (ar[None, None, :] < lengths[:, None, None])
    mask = torch.ones(B, Cc, N, N, dtype=torch.bool) * lm[:, None]
    mask[:, -1] = False
    return mask


# NCHW fp32 score = the layout of the sampling loop, including ragged rows (C*N*N not a multiple of 2048), more samples
# than blocks, and the in-place mode of t2p_pc_run (conditioned positions pre-filled with x_init, never touched);
# 64 x 5 x 128 x 128 is the bench shape
@pytest.mark.parametrize("B,Cc,N,kind", [(2, 5, 32, "random"), (3, 8, 64, "length"), (64, 5, 16, "length"),
                                          (5, 5, 6, "random"), (3, 3, 10, "length"), (2, 5, 128, "none"),
                                          (64, 5, 128, "length"), (300, 5, 32, "length"), (2, 8, 256, "length"),
                                          (1024, 5, 16, "length")])
@pytest.mark.parametrize("in_place", [False, True])
def test_pc_steps_loop_layout(B, Cc, N, kind, in_place):
    if in_place and kind == "none":
        pytest.skip("no mask, nothing conditioned")
    g = torch.Generator().manual_seed(16)
    x = torch.randn(B, Cc, N, N, generator=g) * 10
    h = torch.randn(B, Cc, N, N, generator=g)
    labels = torch.randint(0, 50, (B,), generator=g)
    sigmas = torch.tensor(np.exp(np.linspace(np.log(100.0), np.log(0.01), 50)))
    G = torch.rand(B, generator=g) + 0.1
    if kind == "random":
        mask = torch.rand(B, Cc, N, N, generator=g) > 0.3
    elif kind == "length":
        mask = _length_mask(B, Cc, N, g, N // 3)
    else:
        mask = torch.ones(B, Cc, N, N, dtype=torch.bool)
    x_init = torch.randn(B, Cc, N, N, generator=g)
    seed, E = 77, Cc * N * N
    L = _lib.lib()
    dev = {k: v.cuda() for k, v in dict(x=x, h=h, labels=labels, sigmas=sigmas, G=G, x_init=x_init, mask=mask).items()}
    mask_u8 = dev["mask"].contiguous().view(torch.uint8)

    def noise(stream):
        n = torch.empty(B, Cc, N, N, dtype=torch.float32, device="cuda")
        _lib.check(L.t2p_philox_normal(C.c_uint64(seed), stream, 0, n.numel(), C.c_float(1.0), _lib.ptr(n), _st()))
        return n

    # reference math of _ref_steps on the device in float64 (same formulas; the CPU version is too slow at the bench shape)
    score64 = dev["h"].double() / dev["sigmas"][dev["labels"]][:, None, None, None]
    m = dev["mask"] if kind != "none" else torch.ones_like(dev["mask"])
    grad_norm = torch.norm(score64.reshape(B, -1), dim=-1).mean()
    nc, npred = noise(11), noise(12)
    noise_norm = torch.norm(nc.reshape(B, -1), dim=-1).mean()
    step = ((0.17 * noise_norm / grad_norm) ** 2 * 2 * torch.ones(B, device="cuda")).float()
    xc_ref = dev["x"] + step[:, None, None, None] * score64 + torch.sqrt(step * 2)[:, None, None, None] * nc
    xc_ref = torch.where(m, xc_ref, dev["x_init"].double()).float()
    xm_ref = dev["x"] + (dev["G"][:, None, None, None] ** 2) * score64
    xp_ref = torch.where(m, xm_ref + dev["G"][:, None, None, None] * npred, dev["x_init"].double()).float()
    xm_ref = torch.where(m, xm_ref, dev["x_init"].double()).float()

    ws = torch.empty(L.t2p_corrector_workspace_bytes(B, E) // 8, dtype=torch.float64, device="cuda")
    def args(xbuf, stream):
        a = _lib.StepArgs()
        a.x, a.score, a.score_dtype, a.score_nhwc = xbuf.data_ptr(), dev["h"].data_ptr(), _lib.F32, 0
        a.sigmas, a.labels, a.G = dev["sigmas"].data_ptr(), dev["labels"].data_ptr(), dev["G"].data_ptr()
        a.snr = 0.17
        if kind != "none":
            a.mask, a.x_init = mask_u8.data_ptr(), dev["x_init"].data_ptr()
            a.conditioned_in_place = 1 if in_place else 0
        a.seed, a.stream_id, a.B, a.C, a.HW = seed, stream, B, Cc, N * N
        a.workspace = ws.data_ptr()
        return a

    # in-place mode: the caller has applied the condition once (x, x_mean hold x_init where mask == 0); the
    # conditioned entries of the sentinel buffers must come back untouched
    start = torch.where(m, dev["x"], dev["x_init"]) if in_place else dev["x"]
    xc = start.clone()
    _lib.check(L.t2p_corrector_step(C.byref(args(xc, 11)), _st()))
    xp = start.clone()
    xmn = torch.full_like(xp, float("nan"))
    if in_place:
        xmn = torch.where(m, xmn, dev["x_init"])
    a = args(xp, 12)
    a.x_mean_out = xmn.data_ptr()
    _lib.check(L.t2p_predictor_step(C.byref(a), _st()))
    torch.cuda.synchronize()
    for name, got, ref in (("corrector", xc, xc_ref), ("predictor", xp, xp_ref), ("x_mean", xmn, xm_ref)):
        assert torch.isfinite(got).all()
        assert torch.equal(got[~m], dev["x_init"][~m])              # mask handling is bit-exact
        assert rel_err(got, ref) < 2e-6
        if name != "corrector" and m.any():
            # the predictor update is the reference's float64 expression rounded once to float: the same float except
            # where coef * h (the kernel folds G^2 / sigma into one double) rounds differently in the 53rd bit.  (The
            # corrector's step size hangs on an fp32 norm whose last bit depends on the summation order, also in torch.)
            differ = (got[m] != ref[m]).float().mean().item()
            assert differ < 1e-5, differ
    # run-to-run deterministic (fixed reduction order, no atomics)
    xc2 = start.clone()
    _lib.check(L.t2p_corrector_step(C.byref(args(xc2, 11)), _st()))
    torch.cuda.synchronize()
    assert torch.equal(xc, xc2)


@pytest.mark.parametrize("B,H,W,cin,nout", [(2, 128, 128, 128, 5), (3, 32, 32, 64, 8), (1, 30, 48, 128, 5),
                                            (2, 64, 64, 256, 5), (1, 256, 256, 128, 5), (2, 4, 16, 64, 5),
                                            (1, 16, 32, 128, 8)])
def test_final_conv_fused_matches_torch(B, H, W, cin, nout):
    """GroupNorm-affine + SiLU + 3x3 conv to the map channels in one kernel (padding applied after the activation)."""
    g = torch.Generator(device="cuda").manual_seed(8)
    x = (torch.randn(B, cin, H, W, device="cuda", generator=g) * 2).bfloat16().float()
    scale = 1 + 0.3 * torch.randn(B, cin, device="cuda", generator=g)
    shift = 0.5 * torch.randn(B, cin, device="cuda", generator=g)
    w = (torch.randn(nout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16().float()
    bias = torch.randn(nout, device="cuda", generator=g)
    y = F.silu(x * scale[:, :, None, None] + shift[:, :, None, None]).bfloat16().float()
    ref = F.conv2d(y, w, bias, padding=1)
    X = nhwc(x, torch.bfloat16)
    wp = w.permute(0, 2, 3, 1).contiguous().reshape(nout, -1).bfloat16().contiguous()
    out = torch.full((B, nout, H, W), float("nan"), device="cuda")
    _lib.check(_lib.lib().t2p_final_conv(_lib.ptr(X), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(wp), _lib.ptr(bias),
                                         _lib.ptr(out), B, H, W, cin, nout, _st()))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 1.5e-2
