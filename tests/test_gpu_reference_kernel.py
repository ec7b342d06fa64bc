"""The reference's own ``tests/test_kernel.py``, transcribed test by test (same names, same assertions; the `fieldset` fixture is
`ds_2d_left` restated as arrays, see test_gpu_reference_execute.py)."""

import numpy as np
import pytest

import parcels_b200 as pb
from parcels_b200 import AdvectionRK4, AdvectionRK45, Kernel, KernelWarning, Particle, ParticleSet, Variable
from test_gpu_reference_execute import DoNothing, fieldset  # noqa: F401 -- the fixture

pytestmark = pytest.mark.gpu


def MoveEast(particles, fieldset):
    particles.dx += 0.1


def MoveNorth(particles, fieldset):
    particles.dy += 0.1


def test_unknown_var_in_kernel(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])

    def ErrorKernel(particles, fieldset):
        particles.unknown_varname += 0.2

    with pytest.raises(KeyError, match="'unknown_varname'"):
        pset.execute(ErrorKernel, runtime=np.timedelta64(2, "s"), dt=np.timedelta64(1, "s"))


def test_context_in_kernel(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    fieldset.add_context("fix_lon", -0.5)

    def ContextKernel(particles, fieldset):
        particles.x = fieldset.fix_lon

    pset.execute(ContextKernel, runtime=np.timedelta64(2, "s"), dt=np.timedelta64(1, "s"))
    assert pset.x == -0.5


def test_func_context_in_kernel(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])

    def ContextFunc(x):
        return 2 * x

    fieldset.add_context("func", ContextFunc)

    def FuncContextKernel(particles, fieldset):
        particles.x = fieldset.func(particles.x)

    pset.execute(FuncContextKernel, runtime=np.timedelta64(2, "s"), dt=np.timedelta64(1, "s"))
    assert pset.x == 2.0


def test_kernel_init(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    Kernel(kernels=[AdvectionRK4], pset=pset)


def test_kernel_merging(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    merged_kernel = Kernel(kernels=[AdvectionRK4, MoveEast, MoveNorth], pset=pset)
    assert merged_kernel.funcname == "AdvectionRK4MoveEastMoveNorth"
    assert len(merged_kernel._kernels) == 3
    assert merged_kernel._kernels == [AdvectionRK4, MoveEast, MoveNorth]
    merged_kernel = Kernel(kernels=[MoveEast, MoveNorth, AdvectionRK4], pset=pset)
    assert merged_kernel.funcname == "MoveEastMoveNorthAdvectionRK4"
    assert len(merged_kernel._kernels) == 3
    assert merged_kernel._kernels == [MoveEast, MoveNorth, AdvectionRK4]


def test_kernel_from_list(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    kernels_single = Kernel(kernels=[AdvectionRK4], pset=pset)
    kernels_functions = Kernel(kernels=[AdvectionRK4, MoveEast, MoveNorth], pset=pset)
    assert kernels_single.funcname == "AdvectionRK4"
    assert kernels_functions.funcname == "AdvectionRK4MoveEastMoveNorth"


def test_kernel_from_list_error_checking(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    with pytest.raises(ValueError, match="List of `kernels` should have at least one function."):
        Kernel(kernels=[], pset=pset)
    with pytest.raises(TypeError, match=r"Argument `kernels` should be a function or list of functions.*"):
        Kernel(kernels=[AdvectionRK4, "something else"], pset=pset)
    with pytest.raises(TypeError, match=r".* should be a function or list of functions.*"):
        Kernel(kernels=[Kernel(kernels=[AdvectionRK4], pset=pset), MoveEast, MoveNorth], pset=pset)


def test_RK45Kernel_error_no_next_dt(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])
    with pytest.raises(ValueError, match='ParticleClass requires a "next_dt" for AdvectionRK45 Kernel.'):
        Kernel(kernels=[AdvectionRK45], pset=pset)


def test_rk45_kernel_warnings(fieldset):
    pset = ParticleSet(fieldset=fieldset, pclass=Particle.add_variable(Variable("next_dt", dtype=np.float32, initial=1)), x=[0], y=[0], next_dt=1)
    with pytest.warns(KernelWarning):
        pset.execute(AdvectionRK45, runtime=1, dt=1)


def test_kernel_signature(fieldset):
    pset = ParticleSet(fieldset, x=[0.5], y=[0.5])

    def good_kernel(particles, fieldset):
        pass

    def version_3_kernel(particle, fieldset, time):
        pass

    def version_3_kernel_without_time(particle, fieldset):
        pass

    def kernel_switched_args(fieldset, particle):
        pass

    def kernel_with_forced_kwarg(particles, *, fieldset=0):
        pass

    Kernel(kernels=[good_kernel], pset=pset)
    with pytest.raises(ValueError, match="Kernel function must have 2 parameters, got 3"):
        Kernel(kernels=[version_3_kernel], pset=pset)
    with pytest.raises(ValueError, match="Parameter 'particle' has incorrect name. Expected 'particles', got 'particle'"):
        Kernel(kernels=[version_3_kernel_without_time], pset=pset)
    with pytest.raises(ValueError, match="Parameter 'fieldset' has incorrect name. Expected 'particles', got 'fieldset'"):
        Kernel(kernels=[kernel_switched_args], pset=pset)
    with pytest.raises(ValueError, match="Parameter 'fieldset' has incorrect parameter kind. Expected POSITIONAL_OR_KEYWORD, got KEYWORD_ONLY"):
        Kernel(kernels=[kernel_with_forced_kwarg], pset=pset)


@pytest.mark.parametrize("kernel_type", ["update_lon", "update_dlon"])
def test_execution_order(kernel_type):
    # simple_UV_dataset(dims=(1, 1, 2, 2), mesh="flat") with U = [[0, 1], [2, 3]], lon = [0, 2]
    U = np.array([[0.0, 1.0], [2.0, 3.0]]).reshape(1, 1, 2, 2)
    fieldset = pb.FieldSet.from_arrays(lon=np.array([0.0, 2.0]), lat=np.array([-1e6, 1e6]), depth=np.array([0.0]), U=U, V=np.zeros_like(U), mesh="flat")

    def MoveLon_Update_X(particles, fieldset):
        particles.x += 0.2

    def MoveLon_Update_DX(particles, fieldset):
        particles.dx += 0.2

    def SampleP(particles, fieldset):
        particles.p, _ = fieldset.UV[particles]

    SampleParticle = Particle.add_variable(Variable("p", dtype=np.float32, initial=0.0))
    MoveLon = MoveLon_Update_DX if kernel_type == "update_dlon" else MoveLon_Update_X
    kernels = [MoveLon, SampleP]
    lons, ps = [], []
    for dir in [1, -1]:
        pset = ParticleSet(fieldset, pclass=SampleParticle, x=0, y=0)
        pset.execute(kernels[::dir], runtime=1, dt=1)
        lons.append(pset.x)
        ps.append(pset.p)
    if kernel_type == "update_dlon":
        assert np.isclose(lons[0], lons[1])
        assert np.isclose(ps[0], ps[1])
        assert np.allclose(lons[0], 0.2)
    else:
        assert np.isclose(ps[0] - ps[1], 0.1)
        assert np.allclose(lons[0], 0.2)


@pytest.mark.parametrize("dt", [1e-2, 1e-5, 1e-6, 1e-9])
def test_small_dt(fieldset, dt):
    pset = ParticleSet(fieldset, x=[0], y=[0])
    pset.execute(DoNothing, dt=dt, runtime=dt * 100)
    assert np.allclose([p.t for p in pset], dt * 100)
