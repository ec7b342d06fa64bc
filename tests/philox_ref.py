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
    # float32 Box-Muller like the device (common.cuh wiener_normals); logf / sinf / cosf of NumPy and CUDA agree to 1-2 float32 ulp,
    # so this restatement equals the device's increments to ~1e-6 relative -- the parity tests inject the DEVICE's own
    # increments (device_normals below) into the oracle, this function pins the stream's integer part and its statistics
    f32 = np.float32
    u1 = (r0.astype(f32) + f32(0.5)) * f32(2.0**-32)
    u2 = (r1.astype(f32) + f32(0.5)) * f32(2.0**-32)
    rad = np.sqrt(f32(-2.0) * np.log(u1))
    ang = (f32(2.0) * u2).astype(np.float64) * np.pi
    return (rad * np.cos(ang).astype(f32)).astype(np.float64), (rad * np.sin(ang).astype(f32)).astype(np.float64)


_ENGINE = {}


def device_normals(seed, rng_call, it, particle_id, device=0):
    """The engine's own Wiener increments for (seed; call, iteration, particle ids) -- pb_debug_normals runs the very device
    function the kernels call.  Feeding these to the oracle pins the DETERMINISTIC part of the diffusion kernels bit for bit."""
    from parcels_b200.engine import Engine

    if device not in _ENGINE:
        _ENGINE[device] = Engine(device)
    pid = np.asarray(particle_id, dtype=np.int64)
    if pid.size == 0:
        return np.zeros(0), np.zeros(0)
    out = _ENGINE[device].debug_normals(seed=seed, rng_call=rng_call, it=it, particle_id=pid)
    return out[:, 0].copy(), out[:, 1].copy()
