"""Particle SoA layout: the default ``Particle`` of the reference (``_core/particle.py:123-222``):
t f64; z, y, x, dz, dy, dx f32; particle_id i64; dt f64; state i32; plus ``ei`` i32 (N, ngrids)."""

from __future__ import annotations

import numpy as np

from .statuscodes import StatusCode

__all__ = ["Particle", "create_particle_data"]


class _ParticleClass:
    variables = (
        ("t", np.float64), ("z", np.float32), ("y", np.float32), ("x", np.float32),
        ("dz", np.float32), ("dy", np.float32), ("dx", np.float32),
        ("particle_id", np.int64), ("dt", np.float64), ("state", np.int32),
    )  # fmt: skip

    def __repr__(self):
        return "Particle(" + ", ".join(f"{n}:{np.dtype(d).name}" for n, d in self.variables) + ")"


Particle = _ParticleClass()


def create_particle_data(*, nparticles, ngrids, initial):
    dtypes = dict(Particle.variables)
    data = {"ei": np.zeros((nparticles, ngrids), dtype=np.int32)}
    for k, v in initial.items():
        v = np.asarray(v)
        if v.shape != (nparticles,):
            raise ValueError(f"Initial value for {k} must have shape ({nparticles},). Got {v.shape=}")
        data[k] = np.ascontiguousarray(v.astype(dtypes[k]))
    for name, dt in Particle.variables:
        if name not in data:
            init = {"dt": 1.0, "state": StatusCode.Evaluate}.get(name, 0)
            data[name] = np.full((nparticles,), init, dtype=dt)
    return data
