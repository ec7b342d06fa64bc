"""Diagnostic (GPU box): per advection-diffusion case, where the device and the oracle (same Philox normals) differ."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from advdiff_run import run_engine_advdiff, run_oracle_advdiff
from engine_run import ulp_diff_f32
from oracle.make_golden import ADVDIFF_CASES
from philox_ref import device_normals, wiener_normals

for name in sys.argv[1:] or list(ADVDIFF_CASES):
    seed = 4242
    ps, err = run_engine_advdiff(name, seed=seed)
    st = {"it": 0}

    def normal(view):
        zx, zy = device_normals(seed, 1, st["it"], view.particle_id)
        st["it"] += 1
        return zx, zy

    pd, oerr = run_oracle_advdiff(name, normal=normal)
    d = ps._data
    print(f"== {name}: err={err!r} oracle_err={oerr} n={len(d['x'])} vs {len(pd['x'])} report={ps.last_report}")
    if len(d["x"]) != len(pd["x"]):
        print("   surviving sets differ:", np.setxor1d(d["particle_id"], pd["particle_id"])[:20])
        continue
    for k in ("particle_id", "state", "t", "dt", "ei"):
        bad = np.flatnonzero(d[k].reshape(len(d["x"]), -1) != pd[k].reshape(len(d["x"]), -1))
        print(f"   {k}: {bad.size} mismatches", (d[k].ravel()[bad[:5]], pd[k].ravel()[bad[:5]]) if bad.size else "")
    for k in "xyz":
        raw = ulp_diff_f32(d[k], pd[k])
        fl = ulp_diff_f32(d[k], pd[k], floor=0.05)
        i = int(np.argmax(fl))
        print(f"   {k}: raw ulp max {raw.max()}  floored max {fl.max():.2f} at i={i}: dev={d[k][i]!r} oracle={pd[k][i]!r} "
              f"absdiff={abs(float(d[k][i]) - float(pd[k][i])):.3e}; floored>1: {(fl > 1).sum()} >4: {(fl > 4).sum()}; top5 {np.sort(fl)[-5:]}")
