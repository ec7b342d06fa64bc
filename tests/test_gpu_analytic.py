"""GPU: the product against the reference's analytic-truth tests (moving / decaying eddy, Brownian std)
and its spatial-hash known answers, through the C-ABI."""

import numpy as np
import pytest

import analytic as A
import parcels_b200 as pb

pytestmark = pytest.mark.gpu
RTOL = {"AdvectionEE": 1e-2, "AdvectionRK2": 1e-4, "AdvectionRK2_3D": 1e-4, "AdvectionRK4": 1e-5, "AdvectionRK4_3D": 1e-5}


@pytest.mark.parametrize("name", list(RTOL))
def test_moving_eddy(name):
    f = A.moving_eddy()
    three_d = name.endswith("_3D")
    fs = pb.FieldSet.from_arrays(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["time"], U=f["U"], V=f["V"],
                                 W=f["V"] if three_d else None, mesh="flat")  # fmt: skip
    ps = pb.ParticleSet(fs, x=[12000.0], y=[12500.0], z=[12500.0], t=[0.0])
    ps.execute(getattr(pb, name), dt=np.timedelta64(30, "m"), endtime=np.timedelta64(1, "h"))
    ex, ey = A.moving_eddy_truth(12000.0, 12500.0, 3600.0)
    np.testing.assert_allclose(ps.x, ex, rtol=RTOL[name])
    np.testing.assert_allclose(ps.y, ey, rtol=RTOL[name])
    if name == "AdvectionRK4_3D":
        np.testing.assert_allclose(ps.z, ey, rtol=RTOL[name])
    assert ps.state[0] == pb.StatusCode.EndofLoop


@pytest.mark.parametrize("name, rtol", [("AdvectionEE", 1e-1), ("AdvectionRK2", 3e-3), ("AdvectionRK4", 1e-5)])
def test_decaying_moving_eddy(name, rtol):
    f = A.decaying_eddy()
    fs = pb.FieldSet.from_arrays(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["time"], U=f["U"], V=f["V"], mesh="flat")
    ps = pb.ParticleSet(fs, x=[10000.0], y=[10000.0], z=[0.0], t=[0.0])
    ps.execute(getattr(pb, name), dt=3600.0, endtime=23 * 3600.0)
    ex, ey = A.decaying_eddy_truth(10000.0, 10000.0, 23 * 3600.0)
    np.testing.assert_allclose(ps.x, ex, rtol=rtol)
    np.testing.assert_allclose(ps.y, ey, rtol=rtol)


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_brownian_std(mesh):
    """reference tests/test_diffusion.py:19-46 with the engine's Philox stream; N raised to 20000 so that the
    same 500 m tolerance is a sharp test (sigma of the std estimate ~ 6 m)."""
    conv = 1 / 1852.0 / 60 if mesh == "spherical" else 1
    lon = np.array([-1e6, 1e6])
    z = np.zeros((1, 1, 2, 2), dtype=np.float32)
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lon.copy(), U=z, V=z, mesh=mesh)
    fs.add_constant_field("Kh_zonal", 100, mesh=mesh)
    fs.add_constant_field("Kh_meridional", 50, mesh=mesh)
    n = 20000
    ps = pb.ParticleSet(fs, x=np.zeros(n), y=np.zeros(n), seed=1234)
    ps.execute(pb.DiffusionUniformKh, runtime=np.timedelta64(2, "h"), dt=np.timedelta64(1, "h"))
    tol = 500 * conv
    np.testing.assert_allclose(np.std(ps.x), np.sqrt(2 * 100 * conv**2 * 7200), atol=tol / 10)
    np.testing.assert_allclose(np.std(ps.y), np.sqrt(2 * 50 * conv**2 * 7200), atol=tol / 10)
    np.testing.assert_allclose(np.mean(ps.x), 0, atol=tol / 10)
    np.testing.assert_allclose(np.mean(ps.y), 0, atol=tol / 10)
    assert (ps.ei[:, -1] == 0).all() and ps.ei.shape[1] == 2  # constant-field evals overwrite ei[:, -1]


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
def test_device_hash_query_known_answers(mesh):
    """reference tests/test_spatialhash.py:50-122 on the device: cell centres resolve to their own cell without a
    hint, NaN / far points give GRID_SEARCH_ERROR (state 52)."""
    lon, lat = A.rotated_grid()
    z = np.zeros((1, 1) + lon.shape, dtype=np.float32)
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, U=z, V=z, mesh=mesh, interp_method="cgrid_velocity")
    eng = fs.engine(0)
    clat, clon, jj, ii = A.cell_centers(lon, lat)
    for f32 in (False, True):
        u, v, w, ei, st = eng.sample_velocity(0.0, 0.0, clat, clon, three_d=False, positions_are_f32=f32, no_hint=True)
        np.testing.assert_array_equal(ei, jj * (lon.shape[1] - 1) + ii)
        assert (st == pb.StatusCode.Evaluate).all()
    if mesh == "flat":
        u, v, w, ei, st = eng.sample_velocity(0.0, 0.0, [lat.mean()], [lon.mean()], three_d=False, no_hint=True)
        assert ei[0] == 29 * 29 + 14
    far = ([-60.0, 80.0], [120.0, -150.0]) if mesh == "spherical" else ([1e5, -1e5], [1e5, -1e5])
    u, v, w, ei, st = eng.sample_velocity(0.0, 0.0, far[0], far[1], three_d=False, no_hint=True)
    assert (st == pb.StatusCode.ErrorGridSearching).all() and (ei == -3 * 29 - 3).all() and (u == 0).all()
    u, v, w, ei, st = eng.sample_velocity(0.0, 0.0, [np.nan, np.inf], [np.nan, np.inf], three_d=False, no_hint=True)
    assert (st == pb.StatusCode.ErrorGridSearching).all()
