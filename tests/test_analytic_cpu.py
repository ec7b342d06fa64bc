"""CPU: the oracle against the reference's analytic-truth and spatial-hash known-answer tests
(reference tests/test_advection.py:254-351, tests/test_spatialhash.py:25-122, tests/test_diffusion.py:19-46)."""

import numpy as np
import pytest

import analytic as A
from oracle import curvilinear_oracle as co
from oracle import parcels_oracle as po

KERN = {"AdvectionEE": (po.AdvectionEE, 1e-2), "AdvectionRK2": (po.AdvectionRK2, 1e-4), "AdvectionRK2_3D": (po.AdvectionRK2_3D, 1e-4),
        "AdvectionRK4": (po.AdvectionRK4, 1e-5), "AdvectionRK4_3D": (po.AdvectionRK4_3D, 1e-5)}  # fmt: skip


@pytest.mark.parametrize("name", list(KERN))
def test_moving_eddy(name):
    f = A.moving_eddy()
    k, rtol = KERN[name]
    three_d = name.endswith("_3D")
    fs = po.OFieldSet(po.OGrid(f["lon"], f["lat"], f["depth"], mesh="flat"), f["U"], f["V"], f["V"] if three_d else None, time=f["time"])
    pd = po.create_particle_data([12000.0], [12500.0], [12500.0], 0.0)
    po.pset_execute(pd, fs, k, 1800.0, endtime=3600.0)
    ex, ey = A.moving_eddy_truth(12000.0, 12500.0, 3600.0)
    np.testing.assert_allclose(pd["x"], ex, rtol=rtol)
    np.testing.assert_allclose(pd["y"], ey, rtol=rtol)
    if name == "AdvectionRK4_3D":
        np.testing.assert_allclose(pd["z"], ey, rtol=rtol)


@pytest.mark.parametrize("name, rtol", [("AdvectionEE", 1e-1), ("AdvectionRK2", 3e-3), ("AdvectionRK4", 1e-5)])
def test_decaying_moving_eddy(name, rtol):
    f = A.decaying_eddy()
    fs = po.OFieldSet(po.OGrid(f["lon"], f["lat"], f["depth"], mesh="flat"), f["U"], f["V"], time=f["time"])
    pd = po.create_particle_data([10000.0], [10000.0], [0.0], 0.0)
    po.pset_execute(pd, fs, KERN[name][0], 3600.0, endtime=23 * 3600.0)
    ex, ey = A.decaying_eddy_truth(10000.0, 10000.0, 23 * 3600.0)
    np.testing.assert_allclose(pd["x"], ex, rtol=rtol)
    np.testing.assert_allclose(pd["y"], ey, rtol=rtol)


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_spatial_hash_known_answers(mesh):
    lon, lat = A.rotated_grid()
    g = po.OGrid(lon, lat, None, mesh=mesh)
    h = co.get_hash(g)
    if mesh == "flat":  # golden statistics of the reference's describe() (tests/test_spatialhash.py:25-47)
        assert (h.bitwidth, h.keys.size, h.faces.size) == (1023, 796054, 1080194)
        assert (h.counts.min(), h.counts.max()) == (1, 4)
        j, i, _ = h.query(np.array([lat.mean(), np.nan]), np.array([lon.mean(), np.nan]))  # :111-122
        assert (j[0], i[0], j[1], i[1]) == (29, 14, -3, -3)
        j, i, c = h.query(np.array([np.nan, np.inf]), np.array([np.nan, np.inf]))  # :50-56
        assert np.all(j == -3) and np.all(i == -3) and np.all(c == -1.0)
    else:
        assert h.bitwidth < 1023  # the entry budget caps the resolution on this tilted mesh (:90-108)
        assert h.faces.size <= max(16 * h.xlow.size, 2**22)
        j, i, _ = h.query(np.array([-60.0, 80.0]), np.array([120.0, -150.0]))  # far outside the regional domain (:83-87)
        assert np.all(j == -3) and np.all(i == -3)
    clat, clon, jj, ii = A.cell_centers(lon, lat)  # every cell centre resolves to its own cell (:59-82)
    j, i, _ = h.query(clat, clon)
    np.testing.assert_array_equal(j, jj)
    np.testing.assert_array_equal(i, ii)


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_nan_node_invalidates_touching_faces(mesh):
    """reference tests/test_spatialhash.py:125-181: the four faces around a NaN node are not in the table, queries at their centres
    give -3, every other cell centre still resolves to its own cell.  Oracle table, and the table of the product's host-side
    build (parcels_b200/spatialhash.py, the one the device table is compared with) -- entry for entry the same."""
    from parcels_b200.spatialhash import build_spatial_hash

    lon, lat = A.rotated_grid()
    clat, clon, jj, ii = A.cell_centers(lon, lat)
    lon, lat = lon.copy(), lat.copy()
    nj, ni = 10, 10
    lon[nj, ni] = lat[nj, ni] = np.nan
    nfx = lon.shape[1] - 1
    touching = [(nj - 1, ni - 1), (nj - 1, ni), (nj, ni - 1), (nj, ni)]
    invalid = {j * nfx + i for j, i in touching}
    h = co.get_hash(po.OGrid(lon, lat, None, mesh=mesh))
    in_table = set(np.unique(h.faces).tolist())
    assert invalid.isdisjoint(in_table) and len(in_table) == jj.size - 4
    j, i, _ = h.query(clat, clon)
    bad = np.isin(jj * nfx + ii, list(invalid))
    assert np.all(j[bad] == -3) and np.all(i[bad] == -3)
    np.testing.assert_array_equal(j[~bad], jj[~bad])
    np.testing.assert_array_equal(i[~bad], ii[~bad])
    t = build_spatial_hash(lon, lat, mesh == "spherical", table=True)
    assert t["bitwidth"] == h.bitwidth and t["n_entries"] == h.faces.size
    np.testing.assert_array_equal(t["faces"], h.faces)
    np.testing.assert_array_equal(t["keys"], h.keys)
    np.testing.assert_array_equal(t["counts"], h.counts)


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_brownian_std(mesh):
    """reference tests/test_diffusion.py:19-46 (DiffusionUniformKh alone, Kh 100/50, 2 h, dt 1 h, N=100, tol 500 m)."""
    conv = 1 / 1852.0 / 60 if mesh == "spherical" else 1
    lon = np.array([-1e6, 1e6])
    z = np.zeros((1, 1, 2, 2), dtype=np.float32)
    fs = po.OFieldSet(po.OGrid(lon, lon.copy(), None, mesh=mesh), z, z, constants={"Kh_zonal": 100, "Kh_meridional": 50})
    pd = po.create_particle_data(np.zeros(100), np.zeros(100), np.zeros(100), 0.0, ngrids=2)
    np.random.seed(1234)
    po.pset_execute(pd, fs, [po.DiffusionUniformKh()], 3600.0, runtime=7200.0)
    tol = 500 * conv
    np.testing.assert_allclose(np.std(pd["x"]), np.sqrt(2 * 100 * conv**2 * 7200), atol=tol)
    np.testing.assert_allclose(np.std(pd["y"]), np.sqrt(2 * 50 * conv**2 * 7200), atol=tol)
    np.testing.assert_allclose(np.mean(pd["x"]), 0, atol=tol)
    np.testing.assert_allclose(np.mean(pd["y"]), 0, atol=tol)


@pytest.mark.parametrize("name", ["AdvectionRK2", "AdvectionRK4", "AdvectionRK45"])
def test_stommel_gyre_conserves_sea_surface_height(name):
    """reference tests/test_advection.py:354-387 (A-grid): P sampled along the trajectory stays within rtol 0.1 of its start
    value over one day of 30-minute steps."""
    f = A.stommel_gyre()
    fs = po.OFieldSet(po.OGrid(f["lon"], f["lat"], None, mesh="flat"), f["U"], f["V"])
    x0 = np.linspace(10e3, 100e3, 2)
    pd = po.create_particle_data(x0, np.full(2, 5000e3), np.zeros(2), 0.0)
    if name == "AdvectionRK45":
        fs.context.update(RK45_tol=0.1, RK45_min_dt=1, RK45_max_dt=24 * 60 * 60)
        pd["next_dt"] = np.zeros(2, dtype=np.float32)
        kern = po.AdvectionRK45
    else:
        kern = KERN[name][0]
    p0 = po.eval_scalar(fs, f["P"], "linear", pd["t"], pd["z"], pd["y"], pd["x"], None).copy()
    po.pset_execute(pd, fs, [kern], 1800.0, runtime=86400.0)
    p1 = po.eval_scalar(fs, f["P"], "linear", pd["t"], pd["z"], pd["y"], pd["x"], None)
    assert np.all(pd["t"] == 86400.0) and np.all(np.abs(pd["x"] - x0) + np.abs(pd["y"] - 5000e3) > 1e3)
    np.testing.assert_allclose(p1, p0, rtol=0.1)
