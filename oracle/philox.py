"""ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the counter-based RNG the CUDA kernels use (cv_ssl_mis_b200/csrc/common.cuh:
philox4x32_10, dropout_keep4; optim.cu: noise_kernel).  This is our own RNG, not the reference's
(torch's Philox stream cannot be reproduced bit-for-bit, SURVEY.md "Hard parts"): it exists so tests can
(a) check the device generator against an independent implementation and (b) inject the very same
dropout masks / noise into the reference-restating oracle.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed: int, stream: int, counters: np.ndarray):
    """counters: uint64 array -> four uint32 arrays."""
    seed &= 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    ctr = counters.astype(np.uint64)
    c0 = ctr & MASK32
    c1 = ctr >> np.uint64(32)
    c2 = np.full_like(c0, stream & 0xFFFFFFFF)
    c3 = np.full_like(c0, 0x5151B200)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _unit(u32):
    return (u32 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def keep_mask(seed, stream, M, C, p, drop_mode, spatial=1):
    """[M, C] float32 keep-mask (1/0) in channels-last order, as b200_dropout_mask writes it."""
    idx = np.arange(M * C, dtype=np.uint64)
    if drop_mode == 2:
        m, c = idx // np.uint64(C), idx % np.uint64(C)
        idx = (m // np.uint64(spatial)) * np.uint64(C) + c
    words = philox4x32_10(seed, stream, idx >> np.uint64(2))
    lane = (idx & np.uint64(3)).astype(np.int64)
    u = np.choose(lane, [_unit(w) for w in words])
    return (u >= np.float32(p)).astype(np.float32).reshape(M, C)


def clamp_noise(seed, stream, n, sigma=0.1, clip=0.2):
    """clamp(sigma * N(0,1), -clip, clip) for n elements (Box-Muller on the four words of counter e // 4)."""
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    x, y, z, w = philox4x32_10(seed, stream, q)
    u0, u1 = np.float32(1.0) - _unit(x), _unit(y)
    u2, u3 = np.float32(1.0) - _unit(z), _unit(w)
    r0 = np.sqrt(np.float32(-2.0) * np.log(u0)).astype(np.float32)
    r1 = np.sqrt(np.float32(-2.0) * np.log(u2)).astype(np.float32)
    two_pi = np.float32(2.0 * np.pi)
    zs = np.stack([r0 * np.cos(two_pi * u1), r0 * np.sin(two_pi * u1), r1 * np.cos(two_pi * u3), r1 * np.sin(two_pi * u3)], 1)
    return np.clip(zs.reshape(-1)[:n].astype(np.float32) * np.float32(sigma), -clip, clip).astype(np.float32)
