"""Particle SoA layout: the default ``Particle`` of the reference (``_core/particle.py:123-222``):
t f64; z, y, x, dz, dy, dx f32; particle_id i64; dt f64; state i32; plus ``ei`` i32 (N, ngrids)."""

from __future__ import annotations

import numpy as np

from .statuscodes import StatusCode

__all__ = ["Particle", "ParticleClass", "Variable", "create_particle_data"]


_CORE = (
    ("t", np.float64), ("z", np.float32), ("y", np.float32), ("x", np.float32),
    ("dz", np.float32), ("dy", np.float32), ("dx", np.float32),
    ("particle_id", np.int64), ("dt", np.float64), ("state", np.int32),
)  # fmt: skip


class Variable:
    """reference _core/particle.py:20-60 (name, dtype, initial)."""

    def __init__(self, name, dtype=np.float32, initial=0, **_ignored):
        self.name, self.dtype, self.initial = name, np.dtype(dtype), initial


class ParticleClass:
    """The default Particle plus optional extra variables (reference ``Particle.add_variable``,
    _core/particle.py:95-113).  Extra variables live in host arrays only: the device kernels never touch
    them, user Python kernels can (they stay aligned through deletions)."""

    def __init__(self, extra=()):
        self.extra = tuple(extra)

    @property
    def variables(self):
        return _CORE + tuple((v.name, v.dtype) for v in self.extra)

    def add_variable(self, variable):
        new = [variable] if isinstance(variable, Variable) else list(variable)
        names = {n for n, _ in self.variables}
        for v in new:
            if v.name in names:
                raise ValueError(f"Variable name already exists: {v.name}")
        return ParticleClass(self.extra + tuple(new))

    def __repr__(self):
        return "Particle(" + ", ".join(f"{n}:{np.dtype(d).name}" for n, d in self.variables) + ")"


Particle = ParticleClass()


def create_particle_data(*, nparticles, ngrids, initial, pclass=None):
    pclass = pclass or Particle
    dtypes = dict(pclass.variables)
    inits = {v.name: v.initial for v in pclass.extra}
    data = {"ei": np.zeros((nparticles, ngrids), dtype=np.int32)}
    for k, v in initial.items():
        v = np.asarray(v)
        if v.shape != (nparticles,):
            raise ValueError(f"Initial value for {k} must have shape ({nparticles},). Got {v.shape=}")
        data[k] = np.ascontiguousarray(v.astype(dtypes[k]))
    for name, dt in pclass.variables:
        if name not in data:
            init = {"dt": 1.0, "state": StatusCode.Evaluate, **inits}.get(name, 0)
            data[name] = np.full((nparticles,), init, dtype=dt)
    return data
