"""NumPy Philox4x32-10 + Box-Muller: independent restatement of the engine's Wiener-increment
stream (csrc/engine.cu: philox4x32_10 / wiener_normals), used to inject identical normals
into the oracle's DiffusionUniformKh."""

from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & MASK).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def wiener_normals(seed, rng_call, it, particle_id):
    pid = np.asarray(particle_id, dtype=np.int64).view(np.uint64)
    n = pid.size
    r0, r1, _, _ = philox4x32_10(
        (pid & MASK).astype(np.uint32), (pid >> np.uint64(32)).astype(np.uint32),
        np.full(n, np.uint32(it & 0xFFFFFFFF)), np.full(n, np.uint32(rng_call & 0xFFFFFFFF)),
        seed & 0xFFFFFFFF, ((seed >> 32) ^ (rng_call >> 32)) & 0xFFFFFFFF,
    )  # fmt: skip
    u1 = (r0.astype(np.float64) + 0.5) * 2.0**-32
    u2 = (r1.astype(np.float64) + 0.5) * 2.0**-32
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 6.283185307179586476925 * u2
    return rad * np.cos(ang), rad * np.sin(ang)
