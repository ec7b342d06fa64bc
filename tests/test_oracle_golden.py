"""CPU tests: pin the oracle (oracle/parcels_oracle.py) against
(1) outputs of the reference's own code stored in tests/golden/ref_cases.npz (bit-exact),
(2) the Parcels-v3 JIT golden trajectories the reference's own regression test uses
    (reference tests/test_interpolation.py:297-378, atol 1e-6),
(3) the reference's known-answer unit tests for this path."""

import os

import numpy as np
import pytest

import cases
from oracle import parcels_oracle as po
from oracle_run import load_case, run_oracle

ERR_NAME = {60: "FieldOutOfBoundError", 61: "FieldOutOfBoundSurfaceError", 70: "OutsideTimeInterval",
            51: "FieldInterpolationError", 52: "GridSearchingError", 50: "GeneralError"}  # fmt: skip


@pytest.fixture(scope="module")
def ref_cases(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_cases.npz"))


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_reference_outputs(name, ref_cases):
    c = load_case(name)
    pd, err = run_oracle(c)
    ref_err = str(ref_cases[f"{name}/error"])
    assert (ERR_NAME.get(err, "") if err else "") == ref_err
    for key in ("particle_id", "state", "ei", "t", "x", "y", "z", "dt"):
        ref = ref_cases[f"{name}/{key}"]
        assert pd[key].shape == ref.shape, key
        np.testing.assert_array_equal(pd[key], ref, err_msg=f"{name}:{key}")


def _v3_case(g):
    """Inputs of the reference regression test: f32 coords, f64 data, flat mesh, A-grid."""
    lon, lat, depth = (g[k].astype(np.float32) for k in ("lon", "lat", "depth"))
    x, y, z = np.meshgrid(np.linspace(0, 1, 7), np.linspace(0, 1, 13), np.linspace(0, 1, 5))
    return lon, lat, depth, x.flatten(), y.flatten(), z.flatten()


V3_FILES = {"linear": "v3_jit_linear.npz", "cgrid_velocity": "v3_jit_cgrid.npz", "freeslip": "v3_jit_freeslip.npz",
            "nearest": "v3_jit_nearest.npz"}  # fmt: skip


@pytest.mark.parametrize("interp", list(V3_FILES))
def test_oracle_reproduces_v3_jit_goldens(golden_dir, interp):
    g = np.load(os.path.join(golden_dir, V3_FILES[interp]))
    lon, lat, depth, x, y, z = _v3_case(g)
    fs = po.OFieldSet(po.OGrid(lon, lat, depth, mesh="flat", offsets=(1, 1, 0)), g["U"], g["V"], g["W"], time=g["time"],
                      interp=interp)  # fmt: skip
    pd = po.create_particle_data(x, y, z, 0.0)
    n = len(x)
    obs = {k: np.full((n, 4), np.nan, dtype=np.float32) for k in "xyz"}

    state = {"i": 0}

    def on_output(pdata, _t):
        i = state["i"]
        if i < 4:
            for k in "xyz":
                obs[k][pdata["particle_id"], i] = pdata[k]
        state["i"] += 1

    po.pset_execute(pd, fs, [po.AdvectionRK4_3D, po.DeleteOnError], 1.0, runtime=4.0, outputdt=1.0, on_output=on_output)
    for k, gk in (("x", "gold_lon"), ("y", "gold_lat"), ("z", "gold_z")):
        np.testing.assert_allclose(obs[k], g[gk], atol=1e-6, equal_nan=True)
    assert np.isnan(g["gold_lon"]).sum() > 0  # deletions are part of the pinned behaviour


def test_search_1d_known_answers():
    """reference tests/test_xgrid.py:242-278."""
    arr = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    idx, b = po.search_1d(arr, np.array([1.1, 2.1, 3.1, 4.5, -0.1, 6.5]))
    np.testing.assert_array_equal(idx, [0, 1, 2, 3, -2, -1])
    np.testing.assert_allclose(b[:4], [0.1, 0.1, 0.1, 0.5])
    # node-exact positions land in the lower cell with bcoord 1
    idx, b = po.search_1d(arr, np.array([1.0, 3.0, 5.0]))
    np.testing.assert_array_equal(idx, [0, 1, 3])
    np.testing.assert_array_equal(b, [0.0, 1.0, 1.0])


def test_xlinear_known_answers():
    """reference tests/test_interpolation.py:33-118: P = x + 2y + 3z + 10*ti on a 3x4x4x4 field."""
    T, Z, Y, X = 3, 4, 4, 4
    ti, zi, yi, xi = np.meshgrid(np.arange(T), np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    P = (xi + 2 * yi + 3 * zi + 10 * ti).astype(np.float64)
    grid = po.OGrid(np.arange(X, dtype=float), np.arange(Y, dtype=float), np.arange(Z, dtype=float))
    time = np.arange(T, dtype=float) * 2.0  # time levels 0, 2, 4 s
    fs = po.OFieldSet(grid, P, P, P, time=time)
    for (t, z, y, x), expected in [((0, 0, 0.49, 0.51), 1.49), ((1, 0, 0.49, 0.51), 6.49), ((1, 2.5, 0.49, 0.51), 13.99)]:
        u, v, w = po.eval_uvw(fs, np.array([float(t)]), np.array([z]), np.array([y]), np.array([x]), None, True)
        np.testing.assert_allclose(u, expected, rtol=1e-14)
        np.testing.assert_allclose(w, expected, rtol=1e-14)


def test_spherical_unit_conversion():
    """reference tests/test_interpolation.py:189-205: u = 1/(1852*60*cos(30 deg)) on a spherical mesh."""
    lon = np.linspace(-10, 10, 5)
    lat = np.linspace(20, 40, 5)
    U = np.ones((1, 1, 5, 5), dtype=np.float32)
    fs = po.OFieldSet(po.OGrid(lon, lat, None, mesh="spherical"), U, U)
    u, v = po.eval_uvw(fs, np.zeros(1), np.zeros(1), np.array([30.0]), np.array([0.0]), None, False)
    np.testing.assert_allclose(u, 1 / (1852 * 60 * np.cos(np.deg2rad(30.0))), rtol=1e-7)
    np.testing.assert_allclose(v, 1 / (1852 * 60), rtol=1e-7)


def test_uniform_flow_displacement():
    """reference tests/test_advection.py:43-61,110-128: 1 m/s for 2 h with dt 15 min."""
    lon = np.linspace(-5, 5, 11)
    lat = np.linspace(40, 50, 11)
    U = np.ones((1, 1, 11, 11), dtype=np.float32)
    fs = po.OFieldSet(po.OGrid(lon, lat, None, mesh="spherical"), U, np.zeros_like(U))
    pd = po.create_particle_data(np.zeros(4), np.array([41.0, 43.0, 45.0, 47.0]), np.zeros(4), 0.0)
    y0 = pd["y"].copy()
    po.pset_execute(pd, fs, po.AdvectionRK4, 900.0, runtime=7200.0)
    np.testing.assert_allclose(pd["x"], 7200 / (1852 * 60 * np.cos(np.deg2rad(y0))), atol=1e-5)
    np.testing.assert_array_equal(pd["state"], po.END_OF_LOOP)


def test_peninsula_streamfunction_conserved():
    """reference tests/test_advection.py:390-425: P conserved along RK4 trajectories, rtol 1e-2."""
    c = load_case("c1_peninsula")
    pd, err = run_oracle(c)
    assert err is None
    g = po.OGrid(c["lon"], c["lat"], None)
    Pf = po.OFieldSet(g, c["P"][None, None], c["P"][None, None])

    def sample(x, y):
        u, _ = po.eval_uvw(Pf, np.zeros(len(x)), np.zeros(len(x)), y, x, None, False)
        return u

    p0 = sample(c["x"].astype(np.float32), c["y"].astype(np.float32))
    p1 = sample(pd["x"], pd["y"])
    np.testing.assert_allclose(p1, p0, rtol=1e-2)
