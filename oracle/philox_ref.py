"""Oracle (test infrastructure): numpy Philox4x32-10 + Box-Muller, the counter-based generator the fused
PC-step kernel uses in place of ``torch.randn_like`` (score_sde_pytorch/sampling.py:164,191;
sde_lib.py:229-230).

Algorithm: Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11), Random123 v1.14
``philox4x32_R(10, ctr, key)``; pinned by the Random123 known-answer vectors in tests/test_philox.py.

Stream layout shared with text2protein_b200/csrc/pc_step.cu (integer part is bit-exact):
  quad q = (global element index) // 4 of the [B_global, C, N, N] tensor;
  counter = (q & 0xffffffff, q >> 32, stream & 0xffffffff, stream >> 32); key = (seed_lo, seed_hi);
  the 4 outputs (x0..x3) give u_i = float32(x_i) * 2^-32 + 2^-33 and
  n0 = r(u0) cos(2 pi u1), n1 = r(u0) sin(2 pi u1), n2 = r(u2) cos(2 pi u3), n3 = r(u2) sin(2 pi u3),
  r(u) = sqrt(-2 ln u), all in float32.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4]; key: uint32 array [..., 2] (broadcastable). Returns uint32 [..., 4]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_bits(seed, stream, first_quad, num_quads):
    q = np.arange(first_quad, first_quad + num_quads, dtype=np.uint64)
    ctr = np.empty((num_quads, 4), dtype=np.uint32)
    ctr[:, 0] = (q & MASK).astype(np.uint32)
    ctr[:, 1] = (q >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(stream & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32((stream >> 32) & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key[None, :])


def bits_to_uniform(x):
    return x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def philox_normal(seed, stream, first_element, count):
    """float32 standard normals for global elements [first_element, first_element + count); both must be
    multiples of 4 (a quad never straddles two samples because C*N*N % 4 == 0)."""
    assert first_element % 4 == 0 and count % 4 == 0
    bits = philox_bits(seed, stream, first_element // 4, count // 4)
    u = bits_to_uniform(bits)
    two_pi = np.float32(6.283185307179586)
    r0 = np.sqrt(np.float32(-2.0) * np.log(u[:, 0]))
    r1 = np.sqrt(np.float32(-2.0) * np.log(u[:, 2]))
    t0 = two_pi * u[:, 1]
    t1 = two_pi * u[:, 3]
    out = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)
    return out.astype(np.float32).reshape(-1)


# stream ids of one sampling run (shared with the kernel and the host sampler)
STREAM_PRIOR = 0


def stream_corrector(i, j, n_steps):
    return 1 + i * (n_steps + 1) + j


def stream_predictor(i, n_steps):
    return 1 + i * (n_steps + 1) + n_steps


# update_fn calls made outside the native loop (generic path: user models / predictors / VPSDE) draw stream
# GENERIC_BASE + n for the n-th call after the run's seed was set (text2protein_b200 sampling._Noise)
GENERIC_BASE = 1 << 40


def stream_generic(n):
    return GENERIC_BASE + n
