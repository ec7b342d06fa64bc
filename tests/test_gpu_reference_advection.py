"""The reference's own ``tests/test_advection.py``, transcribed test by test (same names, same assertions; its xarray datasets
restated as arrays).  Elsewhere already: moving / decaying eddy and Stommel / peninsula (tests/test_gpu_analytic.py,
tests/test_analytic_cpu.py), the v3 goldens (tests/test_gpu_parity.py).  Left out: tests on downloaded NEMO / MITgcm data."""

import numpy as np
import pytest

import parcels_b200 as pb
from parcels_b200 import AdvectionEE, AdvectionRK4, AdvectionRK4_3D, Particle, ParticleSet, StatusCode, Variable

pytestmark = pytest.mark.gpu


def simple_uv(dims=(360, 2, 30, 4), maxdepth=1, mesh="spherical", w=False):
    """`simple_UV_dataset` (reference _datasets/structured/generated.py:10-39) as from_arrays arguments"""
    nt, nz, ny, nx = dims
    max_lon, max_lat = (180.0, 90.0) if mesh == "spherical" else (1e6, 1e6)
    time = np.datetime64("2000-01-01") + (np.arange(nt) * (366 * 86400 / (nt - 1))).astype("timedelta64[s]")
    a = dict(lon=np.linspace(-max_lon, max_lon, nx), lat=np.linspace(-max_lat, max_lat, ny), depth=np.linspace(0, maxdepth, nz), time=time,
             U=np.zeros(dims), V=np.zeros(dims), mesh=mesh)  # fmt: skip
    if w:
        a["W"] = np.zeros(dims)
    return a


@pytest.mark.parametrize("mesh", ["spherical", "flat"])
def test_advection_zonal(mesh, npart=10):
    """Particles at high latitude move geographically faster due to the pole correction."""
    a = simple_uv(mesh=mesh)
    a["U"][:] = 1.0
    fieldset = pb.FieldSet.from_arrays(**a)
    runtime = 7200
    startlat = np.linspace(0, 80, npart)
    startlon = 20.0 + np.zeros(npart)
    pset = ParticleSet(fieldset, x=startlon, y=startlat)
    pset.execute(AdvectionRK4, runtime=runtime, dt=np.timedelta64(15, "m"))
    expected_dlon = runtime
    if mesh == "spherical":
        expected_dlon /= 1852 * 60 * np.cos(np.deg2rad(pset.y))
    np.testing.assert_allclose(pset.x - startlon, expected_dlon, atol=1e-5)
    np.testing.assert_allclose(pset.y, startlat, atol=1e-5)


def test_advection_zonal_with_particlefile(tmp_path):
    npart = 10
    a = simple_uv(mesh="flat")
    a["U"][:] = 1.0
    fieldset = pb.FieldSet.from_arrays(**a)
    pset = ParticleSet(fieldset, x=np.zeros(npart) + 20.0, y=np.linspace(0, 80, npart))
    path = tmp_path / "out.parquet"
    pfile = pb.ParticleFile(path, outputdt=np.timedelta64(30, "m"))
    pset.execute(AdvectionRK4, runtime=np.timedelta64(2, "h"), dt=np.timedelta64(15, "m"), output_file=pfile)
    assert (np.diff(pset.x) < 1.0e-4).all()
    df = pb.read_particlefile(path)
    final = df["t"] == df["t"].max()
    np.testing.assert_allclose(df["x"][final], pset.x, atol=1e-5)


def periodicBC(particles, fieldset):
    particles.total_dlon += particles.dx
    particles.x = np.fmod(particles.x, 2)


def test_advection_zonal_periodic():
    # simple_UV_dataset(dims=(2, 2, 2, 2)) with lon = lat = [0, 2] plus a halo column at lon = 3 (a copy of column 0)
    a = simple_uv(dims=(2, 2, 2, 3), mesh="flat")
    a["U"][:] = 0.1
    a["lon"], a["lat"] = np.array([0.0, 2.0, 3.0]), np.array([0.0, 2.0])
    fieldset = pb.FieldSet.from_arrays(**a)
    PeriodicParticle = Particle.add_variable(Variable("total_dlon", initial=0))
    startlon = np.array([0.5, 0.4])
    pset = ParticleSet(fieldset, pclass=PeriodicParticle, x=startlon, y=[0.5, 0.5])
    pset.execute([AdvectionEE, periodicBC], runtime=np.timedelta64(40, "s"), dt=np.timedelta64(1, "s"))
    np.testing.assert_allclose(pset.total_dlon, 4.0, atol=1e-5)
    np.testing.assert_allclose(pset.x, startlon, atol=1e-5)
    np.testing.assert_allclose(pset.y, 0.5, atol=1e-5)


@pytest.mark.parametrize("mesh", ["spherical", "flat"])
def test_advection_meridional(mesh, npart=10):
    """All particles move the same in meridional direction, regardless of latitude."""
    a = simple_uv(mesh=mesh)
    a["V"][:] = 1.0
    fieldset = pb.FieldSet.from_arrays(**a)
    runtime = 7200
    startlat = np.linspace(0, 80, npart)
    startlon = 20.0 + np.zeros(npart)
    pset = ParticleSet(fieldset, x=startlon, y=startlat)
    pset.execute(AdvectionRK4, runtime=runtime, dt=np.timedelta64(15, "m"))
    expected_dlat = runtime
    if mesh == "spherical":
        expected_dlat /= 1852 * 60
    np.testing.assert_allclose(pset.x, startlon, atol=1e-5)
    np.testing.assert_allclose(pset.y - startlat, expected_dlat, atol=1e-4)


@pytest.mark.parametrize("mesh", ["spherical", "flat"])
def test_horizontal_advection_in_3D_flow(mesh, npart=10):
    """2D zonal flow that increases linearly with z from 0 m/s to 1 m/s."""
    a = simple_uv(mesh=mesh)
    a["U"][:] = 1.0
    a["U"][:, 0, :, :] = 0.0
    fieldset = pb.FieldSet.from_arrays(**a)
    pset = ParticleSet(fieldset, x=np.zeros(npart), y=np.zeros(npart), z=np.linspace(0.1, 0.9, npart))
    pset.execute(AdvectionRK4, runtime=np.timedelta64(2, "h"), dt=np.timedelta64(15, "m"))
    expected_lon = pset.z * pset.t
    if mesh == "spherical":
        expected_lon /= 1852 * 60 * np.cos(np.deg2rad(pset.y))
    np.testing.assert_allclose(pset.x, expected_lon, atol=1.0e-1)


@pytest.mark.parametrize("direction", ["up", "down"])
@pytest.mark.parametrize("resubmerge_particle", [True, False])
def test_advection_3D_outofbounds(direction, resubmerge_particle):
    a = simple_uv(mesh="flat", w=True)
    a["U"][:] = 0.01
    a["W"][:] = -1.0 if direction == "up" else 1.0
    fieldset = pb.FieldSet.from_arrays(**a)

    def DeleteParticle(particles, fieldset):
        particles.state = np.where(particles.state == StatusCode.ErrorOutOfBounds, StatusCode.Delete, particles.state)
        particles.state = np.where(particles.state == StatusCode.ErrorThroughSurface, StatusCode.Delete, particles.state)

    def SubmergeParticle(particles, fieldset):
        if len(particles.state) == 0:
            return
        inds = np.argwhere(particles.state == StatusCode.ErrorThroughSurface).flatten()
        if len(inds) == 0:
            return
        (u, v) = fieldset.UV[particles[inds]]
        particles[inds].dx = u * particles.dt
        particles[inds].dy = v * particles.dt
        particles[inds].dz = 0.0
        particles[inds].z = 0
        particles[inds].state = StatusCode.Evaluate

    kernels = [AdvectionRK4_3D]
    if resubmerge_particle:
        kernels.append(SubmergeParticle)
    kernels.append(DeleteParticle)
    pset = ParticleSet(fieldset=fieldset, x=0.5, y=0.5, z=0.9)
    pset.execute(kernels, runtime=np.timedelta64(10, "s"), dt=np.timedelta64(1, "s"))
    if direction == "up" and resubmerge_particle:
        np.testing.assert_allclose(pset.x[0], 0.6, atol=1e-5)
        np.testing.assert_allclose(pset.z[0], 0, atol=1e-5)
    else:
        assert len(pset) == 0


@pytest.mark.parametrize("u_value, x_slice", [(-0.03, slice(0, 1)), (0.02, slice(None))], ids=["single_u_layer", "full_u"])
@pytest.mark.parametrize("v_value, y_slice", [(0.02, slice(0, 1)), (0.1, slice(None))], ids=["single_v_layer", "full_v"])
@pytest.mark.parametrize("w_value, z_slice", [(None, None), (-0.02, slice(0, 1)), (0.07, slice(None))], ids=["no_vertical", "single_w_layer", "full_w"])
def test_length1dimensions(u_value, x_slice, v_value, y_slice, w_value, z_slice):
    # `ds_2d_padded_high` (reference _datasets/structured/generic.py): T, Z, Y, X = 13, 90, 60, 30; two time levels kept
    T, Z, Y, X = 2, 90, 60, 30
    lon, lat, depth = (2 * np.pi / X * np.arange(X))[x_slice], (2 * np.pi / Y * np.arange(Y))[y_slice], np.arange(Z, dtype=np.float64)
    if w_value:
        depth = depth[z_slice]
    shape = (T, len(depth), len(lat), len(lon))
    time = np.datetime64("2000-01-01") + np.array([0, 30 * 86400 + 43200], dtype="timedelta64[s]")
    fieldset = pb.FieldSet.from_arrays(lon=lon, lat=lat, depth=depth, time=time, U=np.full(shape, u_value), V=np.full(shape, v_value),
                                       W=None if w_value is None else np.full(shape, w_value), mesh="flat")  # fmt: skip
    x0, y0, z0 = 3, 3, 20
    pset = ParticleSet(fieldset, x=x0, y=y0, z=z0)
    kernel = AdvectionRK4 if w_value is None else AdvectionRK4_3D
    pset.execute(kernel, runtime=np.timedelta64(4, "s"), dt=np.timedelta64(1, "s"))
    assert len(pset.x) == len([p.x for p in pset])
    np.testing.assert_allclose(np.array([p.x - x0 for p in pset]), 4 * u_value, atol=1e-5)
    np.testing.assert_allclose(np.array([p.y - y0 for p in pset]), 4 * v_value, atol=1e-5)
    if w_value:
        np.testing.assert_allclose(np.array([p.z - z0 for p in pset]), 4 * w_value, atol=1e-5)


def test_radialrotation(npart=10):
    # `radial_rotation_dataset` (reference _datasets/structured/generated.py:42-88): rigid rotation, period one day
    xdim = ydim = 200
    lon = np.linspace(0, 60, xdim, dtype=np.float32)
    lat = np.linspace(0, 60, ydim, dtype=np.float32)
    omega = 2 * np.pi / 86400.0
    dx, dy = np.meshgrid(lon - np.float32(30.0), lat - np.float32(30.0))
    r, theta = np.sqrt(dx**2 + dy**2), np.arctan2(dy, dx)
    U = np.broadcast_to((r * np.sin(theta) * omega).astype(np.float32), (2, 1, ydim, xdim)).copy()
    V = np.broadcast_to((-r * np.cos(theta) * omega).astype(np.float32), (2, 1, ydim, xdim)).copy()
    fieldset = pb.FieldSet.from_arrays(lon=lon, lat=lat, time=np.array([np.timedelta64(0, "s"), np.timedelta64(10, "D")]), U=U, V=V, mesh="flat")
    dt = np.timedelta64(30, "s")
    lon0 = np.linspace(32, 50, npart)
    lat0 = np.ones(npart) * 30
    starttime = np.arange(np.timedelta64(0, "s"), npart * dt, dt)
    endtime = np.timedelta64(10, "m")
    pset = ParticleSet(fieldset, x=lon0, y=lat0, t=starttime)
    pset.execute(AdvectionRK4, endtime=endtime, dt=dt)
    theta = 2 * np.pi * (pset.t - starttime / np.timedelta64(1, "s")) / (24 * 3600)
    true_lon = (lon0 - 30.0) * np.cos(theta) + 30.0
    true_lat = -(lon0 - 30.0) * np.sin(theta) + 30.0
    np.testing.assert_allclose(pset.x, true_lon, atol=5e-2)
    np.testing.assert_allclose(pset.y, true_lat, atol=5e-2)
