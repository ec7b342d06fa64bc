"""Run a tests/cases.py case through the product (parcels_b200 -> C-ABI -> CUDA)."""

from __future__ import annotations

import numpy as np

import parcels_b200 as pb


def make_fieldset(c):
    fs = pb.FieldSet.from_arrays(lon=c["lon"], lat=c["lat"], depth=c["depth"], time=c["times"], U=c["U"], V=c["V"],
                                 W=c["W"], mesh=c["mesh"], interp_method=c.get("interp", "linear"),
                                 padding=c.get("padding", ("low", "low", "high")))  # fmt: skip
    for k, v in (c["constants"] or {}).items():
        fs.add_constant_field(k, v, mesh=c["mesh"])
    return fs


def run_engine(c, seed=0):
    """Returns (pset, error_class_name or '')."""
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"], seed=seed)
    kern = [getattr(pb, k) for k in c["kernels"]]
    if c["delete_on_error"]:
        kern.append(pb.DeleteParticle)
    err = ""
    try:
        for seg in c["segments"]:
            ps.execute(kern, dt=c["dt"], **seg)
    except RuntimeError as e:
        if type(e).__module__.startswith("parcels_b200._lib"):
            raise
        err = type(e).__name__
    return ps, err


def ulp_diff_f32(a, b, floor=None):
    """Distance in float32 ulps between two float32 arrays (same shape).  ``floor``: measure in units of the float32
    spacing at max(|b|, floor) instead -- for coordinates that cross zero, where the spacing of the value itself says
    nothing about the accuracy of the increments that produced it (floor = the size of one step's displacement)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if floor is not None:
        unit = np.spacing(np.maximum(np.abs(b), np.float32(floor)).astype(np.float32)).astype(np.float64)
        return np.abs(a.astype(np.float64) - b.astype(np.float64)) / unit
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-(2**31)) - ia, ia)
    ib = np.where(ib < 0, np.int64(-(2**31)) - ib, ib)
    return np.abs(ia - ib)
