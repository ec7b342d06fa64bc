"""CPU: the C + OpenMP restatement (bench.py's strong CPU baseline) agrees with the NumPy oracle:
bit-exact on a flat mesh, <= 1 float32 ulp on a spherical mesh (NumPy's float32 cos is its own SIMD routine,
the C port calls glibc cosf)."""

import numpy as np
import pytest

from engine_run import ulp_diff_f32
from oracle import c_port
from oracle import parcels_oracle as po
from oracle_run import load_case


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_c_port_matches_numpy_oracle(mesh):
    c = load_case("c2_small")
    lon, lat = (c["lon"], c["lat"]) if mesh == "spherical" else (c["lon"] * 1e4, (c["lat"] - 30) * 1e4)
    sx = 1.0 if mesh == "spherical" else 1e4
    x, y = c["x"] * sx, (c["y"] if mesh == "spherical" else (c["y"] - 30) * 1e4)
    g = po.OGrid(lon, lat, c["depth"], mesh=mesh)
    fs = po.OFieldSet(g, c["U"], c["V"], c["W"], time=c["times"])
    ref = po.create_particle_data(x, y, c["z"], 0.0)
    po.pset_execute(ref, fs, [po.AdvectionRK4_3D, po.DeleteOnError], 600.0, runtime=7200.0)
    pd = po.create_particle_data(x, y, c["z"], 0.0)
    steps = c_port.advect_rk4_3d(lon=lon, lat=lat, depth=c["depth"], time=c["times"], U=c["U"], V=c["V"], W=c["W"],
                                 spherical=mesh == "spherical", deg2m=g.deg2m, pdata=pd, dt=600.0, endtime=7200.0, threads=4)  # fmt: skip
    assert steps > 0
    for k in ("particle_id", "state", "t", "ei"):
        np.testing.assert_array_equal(pd[k], ref[k], err_msg=k)
    for k in ("x", "y", "z"):
        if mesh == "flat":
            np.testing.assert_array_equal(pd[k], ref[k], err_msg=k)
        else:
            assert ulp_diff_f32(pd[k], ref[k]).max() <= 1, k


def test_cached_reciprocal_division_equals_the_division(tmp_path):
    """csrc/common.cuh div_by_cached (q = a r + two residual corrections, r = RN(1 / b)) is the correctly rounded quotient:
    4e7 cases of oracle/division_check.c (cell-width quotients, random mantissas, the constant divisors 111120 and 6,
    quotients next to rounding midpoints) agree with `/` bit for bit."""
    import os
    import shutil
    import subprocess

    cc = shutil.which(os.environ.get("CC", "gcc"))
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "division_check")
    subprocess.run([cc, "-O2", "-mfma", "-ffp-contract=off", os.path.join(root, "oracle", "division_check.c"), "-o", exe, "-lm"], check=True)
    res = subprocess.run([exe, "40000000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "bad5=0" in res.stdout, res.stdout + res.stderr
