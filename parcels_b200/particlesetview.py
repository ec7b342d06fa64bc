"""Masked, write-through window on the particle SoA that user kernels receive
(reference ``_core/particlesetview.py``: ``ParticleSetView`` / ``ParticleSetViewArray``).

Only USER Python kernels ever see these objects (``stepwise.py``); the built-in kernels run on the
device and never touch them.  Attribute reads give an array-like proxy: arithmetic returns plain
ndarrays (masked copies), in-place operators and item assignment write through to the parent arrays.
"""

from __future__ import annotations

import numpy as np

__all__ = ["ParticleSetView", "ParticleSetViewArray", "SingleParticleView"]


def _global_mask(base: np.ndarray, index) -> np.ndarray:
    """Map an index relative to the view ``base`` (bool mask over all particles) to a global bool mask."""
    out = np.zeros_like(base, dtype=bool)
    if isinstance(index, tuple) and len(index) == 1:
        index = index[0]
    arr = np.asarray(index) if isinstance(index, (np.ndarray, list)) else index
    if isinstance(arr, np.ndarray) and arr.dtype == bool:
        if arr.size == base.size:
            return arr.copy()
        if arr.size == int(base.sum()):
            out[base] = arr
            return out
        raise ValueError(f"Boolean index has incompatible length {arr.size} for selection of size {int(base.sum())}")
    ids = np.flatnonzero(base)
    out[ids[arr]] = True
    return out


class ParticleSetViewArray(np.lib.mixins.NDArrayOperatorsMixin):
    """Proxy of ``data[name][mask]``; every ufunc works on the masked copy, ``out=self`` writes through."""

    __array_priority__ = 100

    def __init__(self, data, index, name):
        self._data, self._index, self._name = data, index, name

    def __array__(self, dtype=None, copy=None):
        arr = self._data[self._name][self._index]
        return arr.astype(dtype) if dtype is not None else arr

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = tuple(i.__array__() if isinstance(i, ParticleSetViewArray) else i for i in inputs)
        if out is not None and any(o is self for o in out):  # in-place operator: compute, then write through
            res = getattr(ufunc, method)(*ins, **kwargs)
            self._data[self._name][self._index] = res
            return self
        if out is not None:
            kwargs["out"] = tuple(o.__array__() if isinstance(o, ParticleSetViewArray) else o for o in out)
        return getattr(ufunc, method)(*ins, **kwargs)

    def __getitem__(self, sub):
        return self.__array__()[sub]

    def __setitem__(self, sub, value):
        if isinstance(sub, tuple) and len(sub) > 1:  # e.g. ei[:, igrid] = ...
            rows = np.flatnonzero(self._index)[sub[0]]
            self._data[self._name][(rows, *sub[1:])] = value
        elif isinstance(sub, slice) and sub == slice(None):
            self._data[self._name][self._index] = value
        else:
            self._data[self._name][_global_mask(self._index, sub)] = value

    def __len__(self):
        return int(np.count_nonzero(self._index))

    def __iter__(self):
        return iter(self.__array__())

    def __repr__(self):
        return repr(self.__array__())

    dtype = property(lambda self: self._data[self._name].dtype)
    shape = property(lambda self: self.__array__().shape)
    size = property(lambda self: int(np.count_nonzero(self._index)))

    def __getattr__(self, attr):  # ndarray methods (min, max, sum, astype, any, ...) on the masked copy
        if attr.startswith("_"):
            raise AttributeError(attr)
        return getattr(self.__array__(), attr)


class ParticleSetView:
    def __init__(self, data, index, fieldset=None):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_index", index)
        object.__setattr__(self, "_fieldset", fieldset)

    def __getattr__(self, name):
        data = object.__getattribute__(self, "_data")
        if name in data:
            return ParticleSetViewArray(data, self._index, name)
        if name.startswith("_"):
            raise AttributeError(name)
        raise KeyError(name)  # like the reference's `self._data[name]` (_core/particlesetview.py:28): an unknown Variable

    def __setattr__(self, name, value):
        if isinstance(value, ParticleSetViewArray):
            value = value.__array__()
        self._data[name][self._index] = value

    def __getitem__(self, index):
        return ParticleSetView(self._data, _global_mask(self._index, index), self._fieldset)

    def __len__(self):
        return len(self._index)  # like the reference: the length of the parent set

    @property
    def size(self):
        return int(np.count_nonzero(self._index))


class SingleParticleView:
    """``pset[i]`` with an integer ``i`` (reference _core/particleset.py:166-168 -> ``ParticleSetView(data, index=i)``):
    attribute reads give the particle's scalar values, assignments write through to the parent arrays."""

    def __init__(self, data, index):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_index", int(index))

    def __getattr__(self, name):
        data = object.__getattribute__(self, "_data")
        if name in data:
            return data[name][self._index]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._data[name][self._index] = value
