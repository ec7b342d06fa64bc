"""Particle SoA layout: the default ``Particle`` of the reference (``_core/particle.py:123-222``):
t f64; z, y, x, dz, dy, dx f32; particle_id i64; dt f64; state i32; plus ``ei`` i32 (N, ngrids)."""

from __future__ import annotations

import operator

import numpy as np

from .statuscodes import StatusCode

__all__ = ["Particle", "ParticleClass", "Variable", "create_particle_data"]


_CORE = (
    ("t", np.float64), ("z", np.float32), ("y", np.float32), ("x", np.float32),
    ("dz", np.float32), ("dy", np.float32), ("dx", np.float32),
    ("particle_id", np.int64), ("dt", np.float64), ("state", np.int32),
)  # fmt: skip


# file attributes of the written core variables (reference _core/particle.py:128-170)
_CORE_ATTRS = {
    "t": {"standard_name": "time", "units": "seconds", "axis": "T"},
    "z": {"standard_name": "vertical coordinate", "units": "m", "positive": "down"},
    "y": {"standard_name": "latitude", "units": "degrees_north", "axis": "Y"},
    "x": {"standard_name": "longitude", "units": "degrees_east", "axis": "X"},
    "particle_id": {"long_name": "Unique identifier for each particle", "cf_role": "trajectory_id"},
}


class Variable:
    """reference _core/particle.py:20-66 (name, dtype, initial, to_write, attrs)."""

    def __init__(self, name, dtype=np.float32, initial=0, to_write=True, attrs=None):
        try:
            dtype = np.dtype(dtype)
        except (TypeError, ValueError) as e:
            raise TypeError(f"Variable dtype must be a valid numpy dtype. Got {dtype=!r}") from e
        if to_write not in (True, False):
            raise ValueError(f"to_write must be one of [True, False]. Got {to_write=!r}")
        attrs = {} if attrs is None else attrs
        if not to_write and attrs != {}:
            raise ValueError(f"Attributes cannot be set if {to_write=!r}.")
        self.name, self.dtype, self.initial, self.to_write, self.attrs = name, dtype, initial, to_write, attrs


class ParticleClass:
    """The default Particle plus optional extra variables (reference ``Particle.add_variable``,
    _core/particle.py:95-113).  Extra variables live in host arrays only: the device kernels never touch
    them, user Python kernels can (they stay aligned through deletions)."""

    def __init__(self, extra=()):
        self.extra = tuple(extra)

    @property
    def variables(self):
        return _CORE + tuple((v.name, v.dtype) for v in self.extra)

    def written_variables(self):
        """Variables with ``to_write`` (reference _core/particlefile.py:193-194), in declaration order."""
        core = [Variable(n, d, attrs=dict(_CORE_ATTRS[n])) for n, d in _CORE if n in _CORE_ATTRS]
        return core + [v for v in self.extra if v.to_write]

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
            if isinstance(init, operator.attrgetter):
                # `Variable("age0", initial=attrgetter("t"))`: a copy of another variable's initial values, in THAT variable's
                # dtype (reference _core/particle.py:212-215)
                data[name] = data[init(_NameOf())].copy()
            else:
                data[name] = np.full((nparticles,), init, dtype=dt)
    return data


class _NameOf:
    """``attrgetter("x")(_NameOf())`` == "x" (reference _compat._attrgetter_helper)."""

    def __getattr__(self, name):
        return name
