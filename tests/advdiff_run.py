"""AdvectionDiffusionM1 / EM cases (oracle/make_golden.py: ADVDIFF_CASES) through the CPU oracle and through the product."""

from __future__ import annotations

import numpy as np

from oracle import parcels_oracle as po
from oracle.make_golden import ADVDIFF_CASES, advdiff_inputs
from oracle_run import load_case, oracle_fieldset


def case_inputs(name):
    base, kern, kd, ktime, dt, runtime, delete, rseed = ADVDIFF_CASES[name]
    c = load_case(base)
    c["W"] = None
    c["z"] = np.abs(np.asarray(c["z"]))
    kz, km, dres = advdiff_inputs(c, kd, ktime)
    return c, kern, kz, km, dres, dt, runtime, delete, rseed


def run_oracle_advdiff(name, normal=None):
    """Returns (pdata, error code or None).  normal=None: the reference's RNG stream (np.random.seed as the fixture)."""
    c, kern, kz, km, dres, dt, runtime, delete, rseed = case_inputs(name)
    ofs = oracle_fieldset(c)
    ofs.scalars = {"Kh_zonal": (kz, "linear"), "Kh_meridional": (km, "linear")}
    ofs.context["dres"] = dres
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    kl = [getattr(po, kern)(normal)] + ([po.DeleteOnError] if delete else [])
    if normal is None:
        np.random.seed(rseed)
    err = None
    try:
        po.pset_execute(pd, ofs, kl, dt, runtime=runtime)
    except po.OracleParticleError as e:
        err = e.code
    return pd, err


def run_engine_advdiff(name, seed=0):
    """Returns (pset, error class name or '')."""
    import parcels_b200 as pb
    from engine_run import make_fieldset

    c, kern, kz, km, dres, dt, runtime, delete, _ = case_inputs(name)
    fs = make_fieldset(c)
    fs.add_field("Kh_zonal", kz)
    fs.add_field("Kh_meridional", km)
    fs.add_context("dres", dres)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"], seed=seed)
    kl = [getattr(pb, kern)] + ([pb.DeleteParticle] if delete else [])
    err = ""
    try:
        ps.execute(kl, dt=dt, runtime=runtime)
    except RuntimeError as e:
        if type(e).__module__.startswith("parcels_b200._lib"):
            raise
        err = type(e).__name__
    return ps, err
