"""The reference's own ``tests/test_mesh.py``, transcribed test by test (same names, same assertions): ``SphericalMesh`` with a
configurable planetary radius, its ``deg2m`` in the grid and in the advection kernel's unit conversion."""

import numpy as np
import pytest

import parcels_b200 as pb
from parcels_b200 import EARTH_RADIUS, AdvectionRK4, ParticleSet, SphericalMesh
from test_gpu_reference_advection import simple_uv

EARTH_DEG2M = EARTH_RADIUS * np.pi / 180


def test_spherical_mesh_deg2m():
    assert SphericalMesh().radius == EARTH_RADIUS
    assert SphericalMesh().deg2m == EARTH_DEG2M
    r = 3389500.0  # Mars radius
    assert SphericalMesh(radius=r).deg2m == pytest.approx(r * np.pi / 180)


@pytest.mark.parametrize("mesh, exp_radius, exp_deg2m", [("spherical", EARTH_RADIUS, EARTH_DEG2M), (SphericalMesh(), EARTH_RADIUS, EARTH_DEG2M),
                                                         (SphericalMesh(radius=3389500.0), 3389500.0, 3389500.0 * np.pi / 180)])  # fmt: skip
def test_xgrid_radius_and_deg2m(mesh, exp_radius, exp_deg2m):
    grid = pb.FieldSet.from_arrays(**dict(simple_uv(), mesh=mesh)).U.grid
    assert grid.is_spherical()
    assert grid.radius == exp_radius
    assert grid.deg2m == pytest.approx(exp_deg2m)


@pytest.mark.gpu
@pytest.mark.parametrize("mesh, deg2m", [(SphericalMesh(), EARTH_DEG2M), (SphericalMesh(radius=3389500.0), 3389500.0 * np.pi / 180),
                                         (SphericalMesh(radius=6051800.0), 6051800.0 * np.pi / 180), (SphericalMesh(radius=EARTH_RADIUS), EARTH_DEG2M)])  # fmt: skip
def test_advection_uses_custom_radius(mesh, deg2m, npart=10):
    a = simple_uv()
    a["U"][:] = 1.0
    fieldset = pb.FieldSet.from_arrays(**dict(a, mesh=mesh))
    runtime = 7200
    startlat = np.linspace(0, 80, npart)
    startlon = 20.0 + np.zeros(npart)
    pset = ParticleSet(fieldset, x=startlon, y=startlat)
    pset.execute(AdvectionRK4, runtime=runtime, dt=np.timedelta64(15, "m"))
    expected_dlon = runtime / (deg2m * np.cos(np.deg2rad(pset.y)))
    np.testing.assert_allclose(pset.x - startlon, expected_dlon, atol=1e-5)
    np.testing.assert_allclose(pset.y, startlat, atol=1e-5)


@pytest.mark.gpu
def test_advection_flat_mesh(npart=10):
    a = simple_uv(mesh="flat")
    a["U"][:] = 1.0
    fieldset = pb.FieldSet.from_arrays(**a)
    runtime = 7200
    startlat = np.linspace(0, 80, npart)
    startlon = 20.0 + np.zeros(npart)
    pset = ParticleSet(fieldset, x=startlon, y=startlat)
    pset.execute(AdvectionRK4, runtime=runtime, dt=np.timedelta64(15, "m"))
    assert fieldset.U.grid.deg2m == 1.0  # flat mesh deg2m
    np.testing.assert_allclose(pset.x - startlon, runtime, atol=1e-5)
    np.testing.assert_allclose(pset.y, startlat, atol=1e-5)


@pytest.mark.parametrize("bad_radius", ["6371000", [6371000], (1, 2), {}])
def test_spherical_mesh_rejects_non_numeric_radius(bad_radius):
    with pytest.raises(TypeError):
        SphericalMesh(radius=bad_radius)


@pytest.mark.parametrize("bad_radius", [0, -1.0, -6371000])
def test_spherical_mesh_rejects_nonpos_radius(bad_radius):
    with pytest.raises(ValueError):
        SphericalMesh(radius=bad_radius)
