"""``parcels_b200.install()`` -- run an UNMODIFIED reference script on the engine.

Patches the reference's seam ``parcels._core.kernel.Kernel.execute(pset, endtime, dt)`` (``_core/kernel.py:174-247``): the
reference's own ``ParticleSet.execute`` keeps its outer loop over output intervals, its argument handling and its
``ParticleFile`` calls (``_core/particleset.py:355-470``); only the inner per-step loop runs on the GPU -- through the same
``ParticleSet._kernel_execute`` the host mirror uses, on the reference's OWN ``pset._data`` arrays (shared, updated in place).

What runs where:
  * built-in kernels of ``parcels.kernels`` (recognised by module + name) and this package's tokens: device kernels;
  * any other Python kernel function in the list (a user ``DeleteParticle``, ageing, periodic boundaries ...): called on the host
    between the device launches of the built-ins, exactly as ``Kernel.execute`` orders them (stepwise.py);
  * a FieldSet the engine has no kernel for (unstructured grids, unsupported interpolators, dask-backed fields ...): the
    ORIGINAL ``Kernel.execute`` is called, unchanged -- never a silent CPU restatement of ours.

Errors: the engine's exceptions carry the reference's class names; they are re-raised as the reference's own classes
(``parcels._core.statuscodes``), so ``except FieldOutOfBoundError`` in the script keeps working.
"""

from __future__ import annotations

import warnings

__all__ = ["install", "uninstall"]

_STATE = {}


def _mirror_fieldset(ref_fieldset):
    fs = getattr(ref_fieldset, "__dict__", {}).get("_b200_fieldset")
    if fs is None:
        from .adapter import from_parcels

        fs = from_parcels(ref_fieldset)
        fs._context_is_reference = True
        try:
            ref_fieldset._b200_fieldset = fs
        except AttributeError:
            pass
    return fs


def _mirror_pset(ref_pset, fs, device):
    ps = getattr(ref_pset, "__dict__", {}).get("_b200_pset")  # (the reference's __getattr__ raises KeyError for unknown names)
    if ps is None or ps.fieldset is not fs or ps._host is not ref_pset._data:
        from .adapter import pset_from_parcels

        ps = pset_from_parcels(ref_pset, fs, device=device)
        object.__setattr__(ref_pset, "_b200_pset", ps)
    return ps


def _translate(exc, ref_codes):
    """The reference's exception class of the same name (statuscodes.py:37-117), same message."""
    cls = getattr(ref_codes, type(exc).__name__, None)
    if isinstance(cls, type) and issubclass(cls, BaseException) and cls is not type(exc):
        return cls(*exc.args)
    return exc


def install(device: int = 0, strict: bool = False):
    """Patch ``parcels._core.kernel.Kernel.execute``.  ``strict=True``: raise instead of falling back to the reference's own
    loop when the engine cannot take a FieldSet / kernel list.  Returns the original method (also kept for ``uninstall``)."""
    import parcels._core.kernel as rk
    import parcels._core.statuscodes as ref_codes

    from . import _lib
    from .particleset import KernelPlan

    if "orig" in _STATE:
        return _STATE["orig"]
    orig = rk.Kernel.execute

    def execute(self, pset, endtime, dt):
        try:
            fs = _mirror_fieldset(self._fieldset)
            for name, value in dict(getattr(self._fieldset, "context", {}) or {}).items():
                fs.context[name] = value  # (the script may add context constants between calls; RK45 tolerances arrive converted)
            ps = _mirror_pset(pset, fs, device)
            plan = KernelPlan(list(self._kernels), fs, ps._pclass)
        except NotImplementedError as e:
            if strict:
                raise
            warnings.warn(f"parcels_b200: this Kernel.execute call stays on the reference's own loop ({e})", RuntimeWarning, stacklevel=2)
            return orig(self, pset, endtime, dt)
        if len(pset._data["x"]) == 0:
            return pset
        ps._host, ps._host_stale = pset._data, False  # the reference owns the arrays between calls (it may have replaced columns)
        ps._device_synced = False
        try:
            ps._kernel_execute(plan, float(endtime), float(dt))
        except _lib.EngineError:
            raise
        except Exception as e:  # noqa: BLE001 -- the mapped particle errors (FieldOutOfBoundError, ...)
            t = _translate(e, ref_codes)
            if t is e:
                raise
            raise t from None
        return pset

    execute.__doc__ = orig.__doc__
    execute._b200_patched = True
    rk.Kernel.execute = execute
    _STATE["orig"], _STATE["module"] = orig, rk
    return orig


def uninstall():
    if "orig" in _STATE:
        _STATE["module"].Kernel.execute = _STATE.pop("orig")
        _STATE.pop("module", None)
