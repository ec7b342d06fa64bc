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


def _assert_str_and_python_varname(name):
    """reference _core/utils/string.py (via _core/particle.py:42)"""
    import keyword

    if not isinstance(name, str):
        raise TypeError(f"Expected a string for variable name, got {type(name).__name__} instead.")
    if not name.isidentifier() or keyword.iskeyword(name):
        raise ValueError(f"Received invalid Python variable name {name!r}: not a valid identifier. HINT: avoid using spaces, special "
                         "characters, and starting with a number.")  # fmt: skip


class Variable:
    """reference _core/particle.py:20-66 (name, dtype, initial, to_write, attrs)."""

    def __init__(self, name, dtype=np.float32, initial=0, to_write=True, attrs=None):
        _assert_str_and_python_varname(name)
        try:
            dtype = np.dtype(dtype)
        except (TypeError, ValueError) as e:
            raise TypeError(f"Variable dtype must be a valid numpy dtype. Got {dtype=!r}") from e
        if to_write not in (True, False):
            raise ValueError(f"to_write must be one of [True, False]. Got {to_write=!r}")
        attrs = {} if attrs is None else attrs
        if not to_write and attrs != {}:
            raise ValueError(f"Attributes cannot be set if {to_write=!r}.")
        self._name, self.dtype, self.initial, self.to_write, self.attrs = name, dtype, initial, to_write, attrs

    name = property(lambda self: self._name)

    def __repr__(self):  # reference _repr_utils.variable_repr
        return f"Variable(name={self.name!r}, dtype={self.dtype!r}, initial={self.initial!r}, to_write={self.to_write!r}, attrs={self.attrs!r})"


class ParticleClass:
    """A class of particles: a list of ``Variable`` objects (reference _core/particle.py:69-113).  The built-in kernels use the
    variables of the default ``Particle``; any further variable lives in host arrays only -- the device kernels never touch it,
    user Python kernels can (it stays aligned through deletions)."""

    def __init__(self, variables):
        if not isinstance(variables, list):
            raise TypeError(f"Expected list of Variable objects, got {type(variables)}")
        if not all(isinstance(var, Variable) for var in variables):
            raise ValueError(f"All items in variables must be instances of Variable. Got {variables=!r}")
        self.variables = variables

    @property
    def extra(self):
        """the variables the default Particle does not have"""
        return tuple(v for v in self.variables if v.name not in _CORE_NAMES)

    def written_variables(self):
        """Variables with ``to_write`` (reference _core/particlefile.py:193-194), in declaration order."""
        return [v for v in self.variables if v.to_write]

    def add_variable(self, variable):
        new = [variable] if isinstance(variable, Variable) else list(variable)
        for v in new:
            if not isinstance(v, Variable):
                raise TypeError(f"Expected Variable, got {type(v)}")
        names = {v.name for v in self.variables}
        for v in new:
            if v.name in names:
                raise ValueError(f"Variable name already exists: {v.name}")
        return ParticleClass(variables=self.variables + new)

    def __repr__(self):  # reference _repr_utils.particleclass_repr
        return "\n".join(repr(v) for v in self.variables)


_CORE_NAMES = {n for n, _ in _CORE}
# the default Particle (reference _core/particle.py:123-178): only the variables with file attributes are written
Particle = ParticleClass(variables=[
    Variable(n, d, initial={"dt": 1.0, "state": StatusCode.Evaluate}.get(n, 0), to_write=n in _CORE_ATTRS, attrs=dict(_CORE_ATTRS.get(n, {})))
    for n, d in _CORE
])  # fmt: skip


def create_particle_data(*, nparticles, ngrids, initial=None, pclass=None):
    """reference _core/particle.py:182-222: one array per Variable of the class (+ ``ei`` (N, ngrids) int32)"""
    pclass = pclass or Particle
    initial = {} if initial is None else initial
    dtypes = {v.name: v.dtype for v in pclass.variables}
    data = {"ei": np.zeros((nparticles, ngrids), dtype=np.int32)}
    for k, v in initial.items():
        v = np.asarray(v)
        if v.shape != (nparticles,):
            raise ValueError(f"Initial value for {k} must have shape ({nparticles},). Got {v.shape=}")
        data[k] = np.ascontiguousarray(v.astype(dtypes[k]))
    for var in pclass.variables:
        if var.name not in data:
            init = var.initial
            if isinstance(init, operator.attrgetter):
                # `Variable("age0", initial=attrgetter("t"))`: a copy of another variable's initial values, in THAT variable's
                # dtype (reference _core/particle.py:212-215)
                data[var.name] = data[init(_NameOf())].copy()
            else:
                data[var.name] = np.full((nparticles,), init, dtype=var.dtype)
    return data


class _NameOf:
    """``attrgetter("x")(_NameOf())`` == "x" (reference _compat._attrgetter_helper)."""

    def __getattr__(self, name):
        return name
