"""Run a tests/cases.py case through the CPU oracle (test helper)."""

from __future__ import annotations

import numpy as np

import cases
from oracle import parcels_oracle as po

_KERNELS = {
    "AdvectionRK4": po.AdvectionRK4,
    "AdvectionRK4_3D": po.AdvectionRK4_3D,
    "AdvectionEE": po.AdvectionEE,
    "AdvectionRK2": po.AdvectionRK2,
    "AdvectionRK2_3D": po.AdvectionRK2_3D,
}


def oracle_fieldset(c):
    pad = c.get("padding", ("low", "low", "high"))
    g = po.OGrid(c["lon"], c["lat"], c["depth"], mesh=c["mesh"], offsets=tuple(int(p == "low") for p in pad))
    return po.OFieldSet(g, c["U"], c["V"], c["W"], time=c["times"], constants=c["constants"],
                        interp=c.get("interp", "linear"))  # fmt: skip


def run_oracle(c, normal=None):
    """Returns (pdata, error_code_or_None)."""
    fs = oracle_fieldset(c)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"], ngrids=fs.ngrids)
    kern = []
    for k in c["kernels"]:
        kern.append(po.DiffusionUniformKh(normal) if k == "DiffusionUniformKh" else _KERNELS[k])
    if c["delete_on_error"]:
        kern.append(po.DeleteOnError)
    if c.get("rng_seed") is not None and normal is None:
        np.random.seed(c["rng_seed"])
    err = None
    try:
        for seg in c["segments"]:
            po.pset_execute(pd, fs, kern, c["dt"], **seg)
    except po.OracleParticleError as e:
        err = e.code
    return pd, err


def load_case(name):
    return cases.build(cases.CASES[name])
