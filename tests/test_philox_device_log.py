"""CPU check of the device-side `-2 ln u` (text2protein_b200/csrc/philox.cuh: neg2_log_uniform), restated in numpy
float32 with the coefficients READ from the header, against float64 on every exponent and on random uniforms.

The device Box-Muller replaces logf by this exponent-split polynomial; the normals must stay within the 3e-6 of the
numpy restatement (oracle/philox_ref.py) that DESIGN.md states, i.e. the radius sqrt(-2 ln u) within ~2e-7."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "text2protein_b200", "csrc", "philox.cuh")
f32 = np.float32


def _coefficients():
    src = open(HDR).read()
    body = src[src.index("neg2_log_uniform"):src.index("sqrt_approx")]
    q0 = float(re.search(r"float q = ([-0-9.e]+)f;", body).group(1))
    rest = [float(x) for x in re.findall(r"q = __fmaf_rn\(q, f, ([-0-9.e]+)f\);", body)]
    assert len(rest) == 6
    return [q0] + rest


def _fma(a, b, c):  # float32 fused multiply-add through float64 (exact product, one rounding)
    return (a.astype(np.float64) * np.float64(b) + np.asarray(c, dtype=np.float64)).astype(f32)


def _neg2_log(u):
    coef = _coefficients()
    ix = u.view(np.int32).astype(np.int64)
    e = ((ix - 0x3F2AAAAB) & 0xFF800000).astype(np.uint32).view(np.int32)
    m = (ix.astype(np.int32) - e).view(f32)
    f = (m - f32(1)).astype(f32)
    k = (e.astype(f32) * f32(1.1920928955078125e-07)).astype(f32)
    q = np.full_like(f, f32(coef[0]))
    for c in coef[1:]:
        q = (q.astype(np.float64) * f.astype(np.float64) + np.float64(f32(c))).astype(f32)
    t = (f * f).astype(f32)
    inner = (q.astype(np.float64) * f.astype(np.float64) - 0.5).astype(f32)
    p = (inner.astype(np.float64) * t.astype(np.float64) + f.astype(np.float64)).astype(f32)
    ln = (k.astype(np.float64) * np.float64(f32(0.6931471805599453)) + p.astype(np.float64)).astype(f32)
    return (ln * f32(-2)).astype(f32)


def _uniform(x):  # philox_uniform of the header / oracle
    return (x.astype(f32) * f32(2.3283064365386963e-10) + f32(1.1641532182693481e-10)).astype(f32)


def test_device_log_matches_float64():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2 ** 32, size=2_000_000, dtype=np.uint64).astype(np.uint32)
    edge = np.array([0, 1, 2, 3, 2 ** 32 - 1, 2 ** 32 - 2, 2 ** 31, 2 ** 31 - 1, 0xAAAAAAAA, 0xAAAAAAAB, 0x55555555],
                    dtype=np.uint32)
    every_exponent = (np.uint64(1) << np.arange(32, dtype=np.uint64)).astype(np.uint32)
    u = _uniform(np.concatenate([x, edge, every_exponent, every_exponent - 1]))
    assert u.min() > 0 and u.max() <= 1
    v = _neg2_log(u)
    ref = -2.0 * np.log(u.astype(np.float64))
    assert (v >= 0).all()                      # sqrt never sees a negative argument
    assert np.all(v[u == 1] == 0)
    radius_err = np.abs(np.sqrt(v.astype(np.float64)) - np.sqrt(ref)).max()
    assert radius_err < 2.5e-7, radius_err
    rel = np.abs(v - ref)[ref > 0] / ref[ref > 0]
    assert rel.max() < 2e-7, rel.max()
