"""The reference's own ``tests/test_particleset_execute.py``, transcribed test by test against this package (same names, same
assertions; the xarray datasets of its fixtures restated as arrays: `ds_2d_left` -> `fieldset`, `simple_UV_dataset` ->
`zonal_flow_fieldset` / `time_varying_zonal_flow_fieldset`).  Left out: the unstructured-grid (`ux*`) tests, the two tests whose
body is `...` in the reference, and `test_errorinterpolation` (a user-defined Python VectorInterpolator: the fallback path, not
this library's)."""

from contextlib import nullcontext as does_not_raise
from datetime import datetime, timedelta

import numpy as np
import pytest

import parcels_b200 as pb
from parcels_b200 import FieldOutOfBoundError, OutsideTimeInterval, Particle, ParticleSet, StatusCode, Variable
from parcels_b200 import AdvectionEE, AdvectionRK2, AdvectionRK4, AdvectionRK45

pytestmark = pytest.mark.gpu

T, Z, Y, X = 13, 90, 60, 30
TIME = np.datetime64("2000-01-01") + (np.arange(T) * (366 * 86400 / (T - 1))).astype("timedelta64[s]")  # xr.date_range("2000", "2001", 13)


def DoNothing(particles, fieldset):
    pass


def DeleteParticle(particles, fieldset):
    particles.state = np.where(particles.state >= 50, StatusCode.Delete, particles.state)


@pytest.fixture
def fieldset():
    """`ds_2d_left` (reference _datasets/structured/generic.py:157-205): U_A_grid / V_A_grid random, flat mesh"""
    rng = np.random.default_rng(0)
    return pb.FieldSet.from_arrays(lon=2 * np.pi / X * np.arange(X), lat=2 * np.pi / Y * np.arange(Y), depth=np.arange(Z, dtype=np.float64),
                                   time=TIME, U=rng.random((T, Z, Y, X)), V=rng.random((T, Z, Y, X)), mesh="flat")  # fmt: skip


@pytest.fixture
def fieldset_no_time_interval():
    rng = np.random.default_rng(0)
    return pb.FieldSet.from_arrays(lon=2 * np.pi / X * np.arange(X), lat=2 * np.pi / Y * np.arange(Y), depth=np.arange(Z, dtype=np.float64),
                                   U=rng.random((1, Z, Y, X)), V=rng.random((1, Z, Y, X)), mesh="flat")  # fmt: skip


def _simple_uv(dims=(360, 2, 30, 4), time=None):
    """`simple_UV_dataset(mesh="flat")` (reference _datasets/structured/generated.py:10-39)"""
    nt, nz, ny, nx = dims
    if time is None:
        time = np.datetime64("2000-01-01") + (np.arange(nt) * (366 * 86400 / (nt - 1))).astype("timedelta64[s]")
    return dict(lon=np.linspace(-1e6, 1e6, nx), lat=np.linspace(-1e6, 1e6, ny), depth=np.linspace(0, 1, nz), time=time,
                U=np.zeros(dims), V=np.zeros(dims), mesh="flat")  # fmt: skip


@pytest.fixture
def zonal_flow_fieldset():
    a = _simple_uv()
    a["U"][:] = 1.0
    return pb.FieldSet.from_arrays(**a)


@pytest.fixture
def time_varying_zonal_flow_fieldset():
    nt = 25
    times = np.array([np.timedelta64(3 * i, "h") for i in range(nt)])
    a = _simple_uv(dims=(nt, 2, 6, 6), time=times)
    u = np.cos(2 * np.pi * (times / np.timedelta64(1, "s")) / 86400.0)
    a["U"][:] = u[:, None, None, None]
    return pb.FieldSet.from_arrays(**a)


def test_execute_trajectory_independent_of_other_particles_release_times(time_varying_zonal_flow_fieldset):
    fieldset = time_varying_zonal_flow_fieldset
    t0 = np.timedelta64(0, "s")

    def run(release_times):
        npart = len(release_times)
        pset = ParticleSet(fieldset, pclass=Particle, t=np.array(release_times), z=np.zeros(npart), y=np.zeros(npart), x=np.zeros(npart))
        pset.execute(AdvectionRK4, dt=np.timedelta64(1, "h"), endtime=np.timedelta64(48, "h"))
        return pset.x[0], pset.y[0]

    alone = run([t0])
    uniform = run([t0] * 4)
    staggered = run([t0] + [t0 + np.timedelta64(3, "h")] * 3)
    assert uniform == pytest.approx(alone)
    assert staggered == pytest.approx(alone)


def test_pset_execute_invalid_arguments(fieldset, fieldset_no_time_interval):
    with pytest.raises(ValueError, match="dt must be a non-zero datetime.timedelta or np.timedelta64 object, got .*"):
        ParticleSet(fieldset, x=[0.2], y=[5.0], pclass=Particle).execute(AdvectionRK4, dt=np.timedelta64(0, "s"))
    with pytest.raises(ValueError, match="runtime and endtime are mutually exclusive - provide one or the other. Got .*"):
        ParticleSet(fieldset, x=[0.2], y=[5.0], pclass=Particle).execute(
            AdvectionRK4, runtime=np.timedelta64(1, "s"), endtime=np.datetime64("2100-01-01"), dt=np.timedelta64(1, "s"))  # fmt: skip
    msg = "Calculated/provided end time of .* is not in fieldset time interval .*"
    with pytest.raises(ValueError, match=msg):
        ParticleSet(fieldset, x=[0.2], y=[5.0], pclass=Particle).execute(AdvectionRK4, endtime=np.datetime64("1990-01-01"), dt=np.timedelta64(1, "s"))
    with pytest.raises(ValueError, match=msg):
        ParticleSet(fieldset, x=[0.2], y=[5.0], pclass=Particle).execute(AdvectionRK4, endtime=np.datetime64("2100-01-01"), dt=np.timedelta64(-1, "s"))
    with pytest.raises(ValueError, match="The endtime must be of the same type as the fieldset.time_interval start time. Got .*"):
        ParticleSet(fieldset, x=[0.2], y=[5.0], pclass=Particle).execute(AdvectionRK4, endtime=12345, dt=np.timedelta64(1, "s"))
    with pytest.raises(ValueError, match="The runtime must be provided when the time_interval is not defined for a fieldset."):
        ParticleSet(fieldset_no_time_interval, x=[0.2], y=[5.0], pclass=Particle).execute(AdvectionRK4, dt=np.timedelta64(1, "s"))


@pytest.mark.parametrize("runtime, expectation", [(np.timedelta64(5, "s"), does_not_raise()), (timedelta(seconds=2), does_not_raise()),
                                                  (5.0, does_not_raise()), (np.datetime64("2001-01-02T00:00:00"), pytest.raises(ValueError)),
                                                  (datetime(2000, 1, 2, 0, 0, 0), pytest.raises(ValueError))])  # fmt: skip
def test_particleset_runtime_type(fieldset, runtime, expectation):
    pset = ParticleSet(fieldset, x=[0.2], y=[5.0], z=[50.0], pclass=Particle)
    with expectation:
        pset.execute(runtime=runtime, dt=np.timedelta64(10, "s"), kernels=DoNothing)


@pytest.mark.parametrize("endtime, expectation", [(np.datetime64("2000-01-02T00:00:00"), does_not_raise()), (5.0, pytest.raises(ValueError)),
                                                  (np.timedelta64(5, "s"), pytest.raises(ValueError)), (timedelta(seconds=2), pytest.raises(ValueError)),
                                                  (datetime(2000, 1, 2, 0, 0, 0), pytest.raises(ValueError))])  # fmt: skip
def test_particleset_endtime_type(fieldset, endtime, expectation):
    pset = ParticleSet(fieldset, x=[0.2], y=[5.0], z=[50.0], pclass=Particle)
    with expectation:
        pset.execute(endtime=endtime, dt=np.timedelta64(10, "m"), kernels=DoNothing)


def test_sampleUonly(fieldset):
    def SampleU(particles, fieldset):
        _ = fieldset.U[particles]

    pset = ParticleSet(fieldset, x=[0.2], y=[5.0])
    with pytest.warns(RuntimeWarning, match="Sampling of velocities should normally be done using fieldset.UV or fieldset.UVW object; tread carefully"):
        pset.execute(SampleU, runtime=np.timedelta64(1, "D"), dt=np.timedelta64(1, "D"))


def test_particleset_run_to_endtime(fieldset):
    def SampleUV(particles, fieldset):
        _, _ = fieldset.UV[particles]

    pset = ParticleSet(fieldset, x=[0.2], y=[5.0], t=[TIME[0]])
    pset.execute(SampleUV, endtime=TIME[-1], dt=np.timedelta64(1, "D"))
    assert np.timedelta64(int(pset[0].t), "s") + TIME[0] == TIME[-1]


@pytest.mark.parametrize("kernel", [AdvectionEE, AdvectionRK2, AdvectionRK4, AdvectionRK45])
@pytest.mark.parametrize("dt", [np.timedelta64(10, "D"), np.timedelta64(1, "D")])
def test_particleset_run_RK_to_endtime_fwd_bwd(kernel, dt):
    """RK kernels can be run to the end of a fieldset's time interval (and do not throw OutsideTimeInterval)"""
    a = _simple_uv(dims=(T, 2, Y, X), time=TIME)  # zero velocities to avoid out-of-bounds errors
    a.update(lon=2 * np.pi / X * np.arange(X), lat=2 * np.pi / Y * np.arange(Y))
    fieldset = pb.FieldSet.from_arrays(**a)
    pclass = Particle
    if kernel is AdvectionRK45:
        fieldset.add_context("RK45_tol", 10)
        fieldset.add_context("RK45_min_dt", 1)
        fieldset.add_context("RK45_max_dt", 24 * 60 * 60)
        pclass = Particle.add_variable(Variable("next_dt"))
    pset = ParticleSet(fieldset, pclass=pclass, x=[0.2], y=[5.0], t=[TIME[0]])
    pset.execute(kernel, endtime=TIME[-1], dt=dt)
    assert pset[0].t == 366 * 86400.0
    pset.execute(kernel, endtime=TIME[0], dt=-dt)
    assert pset[0].t == 0.0


def test_particleset_interpolate_on_domainedge(zonal_flow_fieldset):
    fieldset = zonal_flow_fieldset
    MyParticle = Particle.add_variable(Variable("var"))

    def SampleUV(particles, fieldset):
        particles.var, _ = fieldset.UV[particles]

    pset = ParticleSet(fieldset, pclass=MyParticle, x=fieldset.U.grid.lon[-1], y=fieldset.U.grid.lat[-1])
    pset.execute(SampleUV, runtime=np.timedelta64(1, "D"), dt=np.timedelta64(1, "D"))
    np.testing.assert_equal(pset[0].var, 1)


def test_particleset_interpolate_outside_domainedge(zonal_flow_fieldset):
    fieldset = zonal_flow_fieldset

    def SampleU(particles, fieldset):
        particles.dx, _ = fieldset.UV[particles]

    pset = ParticleSet(fieldset, x=fieldset.U.grid.lon[-1], y=fieldset.U.grid.lat[-1] + 1e-3)
    with pytest.raises(FieldOutOfBoundError):
        pset.execute(SampleU, runtime=np.timedelta64(2, "D"), dt=np.timedelta64(1, "D"))


@pytest.mark.parametrize("dt", [np.timedelta64(1, "s"), np.timedelta64(1, "ms"), np.timedelta64(10, "ms"), np.timedelta64(1, "ns")])
def test_pset_execute_subsecond_dt(fieldset, dt):
    def AddDt(particles, fieldset):
        particles.added_dt += particles.dt

    pclass = Particle.add_variable(Variable("added_dt", dtype=np.float32, initial=0))
    pset = ParticleSet(fieldset, pclass=pclass, x=0, y=0)
    pset.execute(AddDt, runtime=dt * 10, dt=dt)
    np.testing.assert_allclose(pset[0].added_dt, 10.0 * (dt / np.timedelta64(1, "s")), atol=1e-5)


def test_pset_remove_particle_in_kernel(fieldset):
    npart = 100
    pset = ParticleSet(fieldset, x=np.linspace(0, 1, npart), y=np.linspace(1, 0, npart))

    def DeleteKernel(particles, fieldset):
        particles.state = np.where((particles.x >= 0.4) & (particles.x <= 0.6), StatusCode.Delete, particles.state)

    pset.execute(DeleteKernel, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(1, "s"))
    indices = [i for i in range(npart) if not (40 <= i < 60)]
    assert [p.particle_id for p in pset] == indices
    assert pset[70].particle_id == 90
    assert pset[-1].particle_id == npart - 1
    assert pset.size == 80


@pytest.mark.parametrize("npart", [1, 100])
def test_pset_stop_simulation(fieldset, npart):
    pset = ParticleSet(fieldset, x=np.zeros(npart), y=np.zeros(npart), pclass=Particle)

    def Delete(particles, fieldset):
        particles[particles.t >= 4].state = StatusCode.StopExecution

    pset.execute(Delete, dt=np.timedelta64(1, "s"), runtime=np.timedelta64(21, "s"))
    assert pset[0].t == 4


@pytest.mark.parametrize("with_delete", [True, False])
def test_pset_multi_execute(fieldset, with_delete, npart=10, n=5):
    pset = ParticleSet(fieldset, x=np.linspace(0, 1, npart), y=np.zeros(npart))

    def AddLat(particles, fieldset):
        particles.dy += 0.1

    for _ in range(n):
        pset.execute(AddLat, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(1, "s"))
        if with_delete:
            pset.remove_indices(len(pset) - 1)
    if with_delete:
        assert np.allclose(pset.y, n * 0.1, atol=1e-12)
    else:
        assert np.allclose([p.y - n * 0.1 for p in pset], np.zeros(npart), rtol=1e-12)


def test_some_particles_throw_outofbounds(zonal_flow_fieldset):
    npart = 100
    lon = np.linspace(0, 9e5, npart)
    pset = ParticleSet(zonal_flow_fieldset, x=lon, y=np.zeros_like(lon))
    with pytest.raises(FieldOutOfBoundError):
        pset.execute(AdvectionEE, runtime=np.timedelta64(1_000_000, "s"), dt=np.timedelta64(10_000, "s"))


def test_delete_on_all_errors(fieldset):
    def MoveRight(particles, fieldset):
        particles.dx += 1
        fieldset.UV[particles.t, particles.z, particles.y, particles.x, particles]

    def DeleteAllErrorParticles(particles, fieldset):
        particles[particles.state > 20].state = StatusCode.Delete

    pset = ParticleSet(fieldset, x=[1e5, 2], y=[0, 0])
    pset.execute([MoveRight, DeleteAllErrorParticles], runtime=np.timedelta64(10, "s"), dt=np.timedelta64(1, "s"))
    assert len(pset) == 0


def test_some_particles_throw_outoftime(fieldset):
    time = [TIME[0] + np.timedelta64(t, "D") for t in [0, 350]]
    pset = ParticleSet(fieldset, x=np.zeros(2), y=np.zeros(2), t=time)

    def FieldAccessOutsideTime(particles, fieldset):
        fieldset.UV[particles.t + 400 * 86400, particles.z, particles.y, particles.x, particles]

    with pytest.raises(OutsideTimeInterval):
        pset.execute(FieldAccessOutsideTime, runtime=np.timedelta64(1, "D"), dt=np.timedelta64(10, "D"))


def test_execution_check_stopallexecution(fieldset):
    def addoneLon(particles, fieldset):
        particles.dx += 1
        particles[particles.x + particles.dx >= 10].state = StatusCode.StopAllExecution

    pset = ParticleSet(fieldset, x=[0, 0], y=[0, 0])
    pset.execute(addoneLon, runtime=np.timedelta64(20, "s"), dt=np.timedelta64(1, "s"))
    np.testing.assert_allclose(pset.x, 9)
    np.testing.assert_allclose(pset.t, 9)


def test_execution_recover_out_of_bounds(fieldset):
    npart = 2

    def MoveRight(particles, fieldset):
        fieldset.UV[particles.t, particles.z, particles.y, particles.x + 0.1, particles]
        particles.dx += 0.1

    def MoveLeft(particles, fieldset):
        inds = np.where(particles.state == StatusCode.ErrorOutOfBounds)
        particles[inds].dx -= 1.0
        particles[inds].state = StatusCode.Success

    lon = np.linspace(0.05, 6.95, npart)
    lat = np.linspace(1, 0, npart)
    pset = ParticleSet(fieldset, x=lon, y=lat)
    pset.execute([MoveRight, MoveLeft], runtime=np.timedelta64(60, "s"), dt=np.timedelta64(1, "s"))
    assert len(pset) == npart
    np.testing.assert_allclose(pset.x, [6.05, 5.95], rtol=1e-5)
    np.testing.assert_allclose(pset.y, lat, rtol=1e-5)


@pytest.mark.parametrize("npart", [1, 100])
def test_execution_fail_python_exception(fieldset, npart):
    pset = ParticleSet(fieldset, x=np.linspace(0, 1, npart), y=np.linspace(1, 0, npart))

    def PythonFail(particles, fieldset):
        inds = np.argwhere(particles.t >= 10)
        if inds.size > 0:
            raise RuntimeError("Enough is enough!")

    with pytest.raises(RuntimeError):
        pset.execute(PythonFail, runtime=np.timedelta64(20, "s"), dt=np.timedelta64(2, "s"))
    assert len(pset) == npart
    assert all(pset.t == 10)


@pytest.mark.parametrize("kernel_names, expected", [("Lat1", [0, 1]), ("Lat2", [2, 0]), ("Lat1and2", [2, 1]), ("Lat1then2", [2, 1])])
def test_execution_update_particle_in_kernel_function(fieldset, kernel_names, expected):
    npart = 2
    pset = ParticleSet(fieldset, x=np.linspace(0, 1, npart), y=np.zeros(npart))

    def Lat1(particles, fieldset):
        def SetLat1(p):
            p.y = 1

        SetLat1(particles[(particles.y == 0) & (particles.x > 0.5)])

    def Lat2(particles, fieldset):
        def SetLat2(p):
            p.y = 2

        SetLat2(particles[(particles.y == 0) & (particles.x < 0.5)])

    def Lat1and2(particles, fieldset):
        def SetLat1(p):
            p.y = 1

        def SetLat2(p):
            p.y = 2

        SetLat1(particles[(particles.y == 0) & (particles.x > 0.5)])
        SetLat2(particles[(particles.y == 0) & (particles.x < 0.5)])

    kernels = {"Lat1": [Lat1], "Lat2": [Lat2], "Lat1and2": [Lat1and2], "Lat1then2": [Lat1, Lat2]}[kernel_names]
    pset.execute(kernels, runtime=np.timedelta64(2, "s"), dt=np.timedelta64(1, "s"))
    np.testing.assert_allclose(pset.y, expected, rtol=1e-5)
