"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the engine's counter-based noise (csrc/agym_common.cuh):
Philox4x32-10 (Salmon et al., SC'11) + Box-Muller on 24-bit uniforms.  Lets a test reproduce, on the host, the noise a
kernel draws for given counters (e.g. the rsample noise of the Doubly-Robust fit)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def make_key(seed, global_run):
    seed, r = int(seed), int(global_run)
    k0 = (seed & 0xFFFFFFFF) ^ ((r * 0x9E3779B9) & 0xFFFFFFFF)
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ ((r + 0x7F4A7C15) & 0xFFFFFFFF)
    return np.uint32(k0), np.uint32(k1)


def philox4x32_10(c0, c1, c2, c3, key):
    c0, c1, c2, c3 = (np.asarray(c, np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def box_muller(a, b):
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = (b >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def normal_x(c0, c1, c2, c3, key):
    """First of the four normals philox_normal4 returns (the .x component)."""
    w = philox4x32_10(c0, c1, c2, c3, key)
    return box_muller(w[0], w[1])[0]
