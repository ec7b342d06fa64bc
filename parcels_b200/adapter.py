"""Build the engine's host mirror from a live FieldSet / ParticleSet of the reference package.

Reads only public attributes of the reference objects (``fieldset.U.data``, ``field.grid.lon``,
``grid._mesh``, ``grid.xdim`` ...; reference ``_core/field.py``, ``_core/xgrid.py``), so it works with
the real xarray-backed objects and with any duck type exposing the same attributes.  Nothing here
computes: it hands NumPy buffers to ``parcels_b200.FieldSet``.
"""

from __future__ import annotations

import numpy as np

from .fieldset import FieldSet, XGrid
from .particleset import ParticleSet

SUPPORTED_VECTOR_INTERP = {"XLinear_Velocity": "linear", "XFreeslip": "freeslip", "XPartialslip": "partialslip",
                           "CGrid_Velocity": "cgrid_velocity", "XNearest_Velocity": "nearest"}  # fmt: skip


def _padding(g):
    """SGRID face-node padding of the X, Y, Z axes as the strings FieldSet takes; only LOW vs not-LOW matters to the
    interpolators (reference interpolators/_xinterpolators.py:99-109, `_get_offsets_dictionary`)."""
    md = g.sgrid_metadata
    out = [str(getattr(fnp.padding, "value", fnp.padding)).lower() for fnp in md.face_dimensions[:2]]
    vd = getattr(md, "vertical_dimensions", None)
    out.append(str(getattr(vd[0].padding, "value", vd[0].padding)).lower() if vd else "high")
    return tuple(out)


def _values(da):
    return np.asarray(getattr(da, "values", da))


def from_parcels(ref_fieldset) -> FieldSet:
    """reference FieldSet -> parcels_b200.FieldSet: rectilinear A-grids (XLinear_Velocity, XFreeslip, XPartialslip),
    rectilinear and curvilinear C-grids (CGrid_Velocity), scalar fields, constant fields and the context constants."""
    U = ref_fieldset.U
    g = U.grid
    vf = getattr(ref_fieldset, "UVW", None) or ref_fieldset.UV
    interp = type(vf.interp_method).__name__
    if interp not in SUPPORTED_VECTOR_INTERP:
        raise NotImplementedError(f"vector interpolator {interp} is not on the engine (supported: {sorted(SUPPORTED_VECTOR_INTERP)})")
    lon, lat = np.asarray(g.lon), np.asarray(g.lat)
    if lon.ndim != 1 and interp not in ("CGrid_Velocity", "XLinear_Velocity"):
        raise NotImplementedError("curvilinear grids are on the engine with CGrid_Velocity and XLinear_Velocity only")
    axes = list(g.axes)
    depth = np.asarray(g.depth) if "Z" in axes else None
    spherical = g._mesh.is_spherical()
    grid = XGrid(lon, lat, depth, mesh="spherical" if spherical else "flat", radius=getattr(g._mesh, "radius", None),
                 xdim=g.xdim, ydim=g.ydim, zdim=g.zdim if "Z" in axes else None)  # fmt: skip
    time = None
    if U.time_interval is not None:
        time = _values(U.data.time)
    W = getattr(ref_fieldset, "W", None)
    fs = FieldSet(grid, _values(U.data), _values(ref_fieldset.V.data), None if W is None else _values(W.data), time=time,
                  interp_method=SUPPORTED_VECTOR_INTERP[interp], padding=_padding(g))
    scalar = {"XLinear": "linear", "XNearest": "nearest", "CGrid_Tracer": "cgrid_tracer", "XLinearInvdistLandTracer": "linear_invdist_land"}
    for name, f in ref_fieldset.fields.items():
        how = type(getattr(f, "interp_method", None)).__name__
        if name not in ("U", "V", "W") and how in scalar and getattr(f, "grid", None) is g and lon.ndim == 1:
            fs.add_field(name, _values(f.data), interp_method=scalar[how])  # sampled on the device (pb_sample_scalar)
    for name, f in ref_fieldset.fields.items():
        if type(getattr(f, "interp_method", None)).__name__ == "XConstantField":
            fs.add_constant_field(name, float(_values(f.data)[0, 0, 0, 0]), mesh="spherical" if f.grid._mesh.is_spherical() else "flat")
    for name, value in dict(getattr(ref_fieldset, "context", {}) or {}).items():  # fieldset.add_context (_core/fieldset.py:207-222)
        fs.add_context(name, value)
    return fs


def pset_from_parcels(ref_pset, fieldset: FieldSet, **kw) -> ParticleSet:
    """reference ParticleSet -> parcels_b200.ParticleSet sharing the SAME SoA arrays."""
    from .particle import _CORE_NAMES, Particle, Variable

    d = ref_pset._data
    # every key of the shared SoA is a Variable of the set's class: extra variables (Particle.add_variable in the reference)
    # must be known here so that deletions go through remove_indices over ALL keys and the set is not kept device-resident
    # with host-only columns (the compacted-download shortcut only handles the default Particle)
    extra = [Variable(k, v.dtype) for k, v in d.items() if k not in _CORE_NAMES and k != "ei"]
    pclass = Particle.add_variable(extra) if extra else Particle
    ps = ParticleSet(fieldset, pclass, x=d["x"], y=d["y"], z=d["z"], t=d["t"], particle_ids=d["particle_id"], **kw)
    ps._data = d  # share the reference's dict of ndarrays: results are written back in place
    ps.eager_host = True  # ... at the end of every execute(), not on first access
    return ps
