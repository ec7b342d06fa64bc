"""GPU: kernel lists that mix built-in kernels (device) with USER Python kernels (host, ParticleSetView) --
the common idiom of the reference's own tests -- against the all-device path and the oracle."""

import numpy as np
import pytest

import parcels_b200 as pb
from engine_run import make_fieldset, run_engine, ulp_diff_f32
from oracle import parcels_oracle as po
from oracle_run import load_case, oracle_fieldset

pytestmark = pytest.mark.gpu


def UserDelete(particles, fieldset):  # the reference's idiom, tests/test_interpolation.py:357-359
    any_error = particles.state >= 50
    particles[any_error].state = pb.StatusCode.Delete


@pytest.mark.parametrize("name", ["c2_small", "through_surface", "curv_sph_2d", "delayed_partial"])
def test_user_delete_kernel_equals_fused_device_path(name):
    c = load_case(name)
    ref, err = run_engine(c)  # [Advection..., pb.DeleteParticle]: one fused kernel launch per Kernel.execute
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    kern = [getattr(pb, k) for k in c["kernels"]] + [UserDelete]
    for seg in c["segments"]:
        ps.execute(kern, dt=c["dt"], **seg)
    assert ps.last_report["mode"] == "stepwise"
    for k in ("particle_id", "state", "t", "ei", "x", "y", "z", "dx"):
        np.testing.assert_array_equal(ps._data[k], ref._data[k], err_msg=k)


def test_periodic_boundary_user_kernel_matches_oracle():
    """a user kernel that acts every step between advection and the position update (reference
    tests/test_advection.py:82-84 style): zonal wrap-around of the displacement."""
    c = load_case("delayed_partial")
    lo, hi = float(c["lon"][3]), float(c["lon"][-4])

    def Periodic(particles, fieldset):
        xn = particles.x + particles.dx
        particles.dx += np.where(xn > hi, lo - hi, 0.0) + np.where(xn < lo, hi - lo, 0.0)

    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    for seg in c["segments"]:
        ps.execute([pb.AdvectionRK4_3D, Periodic, UserDelete], dt=c["dt"], **seg)
    ofs = oracle_fieldset(c)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    for seg in c["segments"]:
        po.pset_execute(pd, ofs, [po.AdvectionRK4_3D, Periodic, po.DeleteOnError], c["dt"], **seg)
    for k in ("particle_id", "state", "t", "ei"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)
    for k in ("x", "y", "z"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)  # flat mesh: bit-exact


def test_user_kernel_can_sample_the_velocity_field():
    """fieldset.UVW[particles] inside a user kernel: hinted by and writing particles.ei, raising particles.state"""
    c = load_case("c2_small")
    fs = make_fieldset(c)
    ofs = oracle_fieldset(c)
    seen = {}

    def Sample(particles, fieldset):
        u, v, w = fieldset.UVW[particles]
        seen["u"], seen["ei"], seen["state"] = u, np.asarray(particles.ei)[:, -1].copy(), np.asarray(particles.state).copy()
        particles[particles.state >= 50].state = pb.StatusCode.Delete

    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    ps.execute([Sample], dt=c["dt"], runtime=c["dt"])
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    view = po.View(pd, np.ones(len(c["x"]), bool))
    ou, ov, ow = po.eval_uvw(ofs, pd["t"], pd["z"], pd["y"], pd["x"], view, True)
    np.testing.assert_array_equal(seen["ei"], pd["ei"][:, -1])
    np.testing.assert_array_equal(seen["state"], pd["state"])
    np.testing.assert_allclose(seen["u"], ou, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("interp", ["linear", "cgrid_velocity"])
def test_v3_regression_written_like_the_reference_test(golden_dir, interp):
    """reference tests/test_interpolation.py:297-378 transcribed line by line onto this package's API: custom
    particle class with a `pid` variable, user-defined DeleteParticle, output every second -- atol 1e-6 vs v3 JIT."""
    import os

    g = np.load(os.path.join(golden_dir, "v3_jit_linear.npz" if interp == "linear" else "v3_jit_cgrid.npz"))
    lon, lat, depth = (g[k].astype(np.float32) for k in ("lon", "lat", "depth"))
    fieldset = pb.FieldSet.from_arrays(lon=lon, lat=lat, depth=depth, time=g["time"], U=g["U"], V=g["V"], W=g["W"], mesh="flat",
                                       interp_method=interp, padding=("low", "low", "high"))  # fmt: skip
    x, y, z = np.meshgrid(np.linspace(0, 1, 7), np.linspace(0, 1, 13), np.linspace(0, 1, 5))
    TestP = pb.Particle.add_variable(pb.Variable("pid", dtype=np.int32, initial=0))
    pset = pb.ParticleSet(fieldset, pclass=TestP, x=x, y=y, z=z, t=np.zeros(x.size), pid=np.arange(x.size))

    def DeleteParticle(particles, fieldset):
        any_error = particles.state >= 50  # This captures all Errors
        particles[any_error].state = pb.StatusCode.Delete

    obs = {k: np.full((x.size, 5), np.nan, dtype=np.float32) for k in "xyz"}

    class Out:
        outputdt = np.timedelta64(1, "s")
        i = 0

        def write(self, ps, _time):
            for k in "xyz":
                obs[k][ps.pid, self.i] = ps._data[k]
            self.i += 1

    pset.execute([pb.AdvectionRK4_3D, DeleteParticle], runtime=np.timedelta64(4, "s"), dt=np.timedelta64(1, "s"), output_file=Out())
    assert pset.last_report["mode"] == "stepwise"
    for k, gk in (("x", "gold_lon"), ("y", "gold_lat"), ("z", "gold_z")):
        np.testing.assert_allclose(obs[k][:, :4], g[gk], atol=1e-6, equal_nan=True)
