"""Edge cases of ParticleSet.execute on the device, transcribed from the reference's own tests
(tests/test_particleset_execute.py:315-357,445-468, tests/test_advection.py:148-190): empty sets, a single particle,
end-time landing with a dt that does not divide the run, per-particle start times in both time directions, everything
deleted in the first step, dt reset, user kernels mixed in."""

import numpy as np
import pytest

import parcels_b200 as pb

pytestmark = pytest.mark.gpu


def _fieldset(u=1.0, v=0.0, tmax=100.0, mesh="flat"):
    lon, lat = np.linspace(0.0, 1000.0, 21), np.linspace(0.0, 500.0, 11)
    U = np.full((2, 1, 11, 21), u, dtype=np.float32)
    V = np.full((2, 1, 11, 21), v, dtype=np.float32)
    return pb.FieldSet.from_arrays(lon=lon, lat=lat, time=np.array([0.0, tmax]), U=U, V=V, mesh=mesh)


@pytest.mark.parametrize("starttime, endtime, dt", [(0, 10, 1), (0, 10, 3), (2, 16, 3), (20, 10, -1), (20, 0, -2), (5, 15, 1)])
def test_execution_endtime(starttime, endtime, dt):
    """reference test_execution_endtime: the last step is clamped so that t lands exactly on endtime."""
    fs = _fieldset()
    ps = pb.ParticleSet(fs, t=float(starttime), x=100.0, y=100.0)
    ps.execute(pb.AdvectionRK4, endtime=float(endtime), dt=float(dt))
    assert ps.t[0] == float(endtime) and ps.dt[0] == float(dt) and ps.state[0] == pb.StatusCode.EndofLoop
    assert ps.x[0] == np.float32(100.0 + (endtime - starttime))  # u = 1 m/s, flat mesh: exact in float32


@pytest.mark.parametrize("starttime, runtime, dt", [(0, 10, 1), (0, 10, 3), (2, 16, 3), (20, 10, -1), (20, 0, -2), (5, 15, 1)])
@pytest.mark.parametrize("npart", [1, 10])
def test_execution_runtime(starttime, runtime, dt, npart):
    fs = _fieldset()
    ps = pb.ParticleSet(fs, t=float(starttime), x=np.full(npart, 200.0), y=np.full(npart, 100.0))
    ps.execute([pb.AdvectionEE, pb.DeleteParticle], runtime=float(runtime), dt=float(dt))
    assert len(ps) == npart and np.all(np.abs(ps.t - starttime - runtime * np.sign(dt)) < 1e-3)


def test_dont_run_particles_outside_starttime():
    """reference test_dont_run_particles_outside_starttime, forward and backward (u = 1 m/s stands in for `x += 1`)."""
    fs = _fieldset()
    for sign, t0 in ((1, 0.0), (-1, 100.0)):
        starts = np.array([t0 + sign * s for s in (0, 2, 10)])
        end = t0 + sign * 8
        ps = pb.ParticleSet(fs, x=np.full(3, 500.0), y=np.full(3, 100.0), t=starts)
        ps.execute(pb.AdvectionRK4, dt=sign * 1.0, endtime=end)
        np.testing.assert_array_equal(ps.x, np.float32([500 + sign * 8, 500 + sign * 6, 500]))
        np.testing.assert_array_equal(ps.t, [end, end, starts[2]])  # the third particle has not been executed
        np.testing.assert_array_equal(ps.state, [1, 1, 10])


def test_empty_set_and_everything_deleted():
    fs = _fieldset()
    ps = pb.ParticleSet(fs, x=[], y=[])
    ps.execute(pb.AdvectionRK4, runtime=10.0, dt=1.0)  # nothing to do, no error
    assert len(ps) == 0
    ps = pb.ParticleSet(fs, x=np.full(5, 2000.0), y=np.full(5, 100.0))  # all outside the domain
    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], runtime=10.0, dt=1.0)
    assert len(ps) == 0 and ps._data["ei"].shape == (0, 1)
    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], runtime=10.0, dt=1.0)  # executing the emptied set again is a no-op
    ps += pb.ParticleSet(fs, x=100.0, y=100.0)
    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], runtime=10.0, dt=1.0)
    assert len(ps) == 1 and ps.x[0] == np.float32(110.0)
    with pytest.raises(pb.FieldOutOfBoundError):  # without a handler the same start raises, like the reference
        pb.ParticleSet(fs, x=np.full(5, 2000.0), y=np.full(5, 100.0)).execute(pb.AdvectionRK4, runtime=10.0, dt=1.0)


def test_some_particles_throw_outofbounds_then_survivors_continue():
    """reference test_some_particles_throw_outofbounds idea: particles leaving through the eastern edge are deleted, the others go on;
    deletions happen in HBM between output intervals (order preserved)."""
    fs = _fieldset(u=10.0)
    x0 = np.linspace(100.0, 990.0, 90)
    ps = pb.ParticleSet(fs, x=x0, y=np.full(90, 100.0))

    class Out:
        outputdt = 10.0
        sizes = []

        def write(self, pset, t):
            self.sizes.append(len(pset))

    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], runtime=50.0, dt=1.0, output_file=Out())
    keep = x0 + 500.0 <= 1000.0
    assert Out.sizes[0] == 90 and Out.sizes[-1] == int(keep.sum()) and sorted(Out.sizes, reverse=True) == Out.sizes
    np.testing.assert_array_equal(ps.particle_id, np.flatnonzero(keep))
    np.testing.assert_allclose(ps.x, x0[keep] + 500.0, rtol=1e-6)


def test_changing_dt_in_kernel_and_dt_reset():
    """reference test_changing_dt_in_kernel: 3 steps for runtime 5 with dt 2 (the last one clamped), dt restored afterwards;
    a user kernel mixed with a device kernel."""
    fs = _fieldset(u=0.0)
    calls = []

    def KernelCounter(particles, fieldset):
        calls.append(len(particles.x))
        particles.dx += 1

    ps = pb.ParticleSet(fs, x=np.zeros(1) + 10, y=np.zeros(1) + 10)
    ps.execute([pb.AdvectionRK4, KernelCounter], dt=2.0, runtime=5.0)
    assert ps.x[0] == 13 and ps.dt[0] == 2 and ps.t[0] == 5 and len(calls) == 3


@pytest.mark.parametrize("name", ["c2_small", "flat_f32c_f64d", "cgrid_rect_3d", "freeslip_3d", "nearest_3d", "rk4_2d_in_3d", "curv_sph_2d"])
@pytest.mark.parametrize("delete", [True, False])
def test_non_finite_positions_are_flagged_like_the_reference(name, delete):
    """Particles released at NaN / +inf / -inf coordinates: the reference's searches put NaN last, +-inf outside, and the NaN test
    on the interpolated value comes BEFORE the out-of-bounds masking (field.py:288-290) -- e.g. x = -inf gives index -2 (not an
    error by itself) and a NaN value: ErrorInterpolation; a NaN depth poisons the Z-lerp.  Same states, times, cells, survivors and
    raised error as the oracle (found with the host-compiled kernels, oracle/hostsim)."""
    import cases
    from engine_run import run_engine
    from oracle_run import run_oracle

    err_name = {60: "FieldOutOfBoundError", 61: "FieldOutOfBoundSurfaceError", 70: "OutsideTimeInterval", 51: "FieldInterpolationError",
                52: "GridSearchingError", 50: "GeneralError"}  # fmt: skip
    for seed in (0, 3, 5):
        rng = np.random.default_rng(seed)
        c = cases.build(dict(cases.CASES[name], delete=delete, n=60))
        for k, val in zip("xyzxyz", (np.nan, np.inf, -np.inf, -np.inf, -np.inf, np.nan)):
            c[k][rng.integers(0, 60)] = val
        ps, err = run_engine(c)
        pd, oerr = run_oracle(c)
        assert err == (err_name[oerr] if oerr else ""), (seed, err, oerr)
        for key in ("particle_id", "state", "t", "ei"):
            np.testing.assert_array_equal(ps._data[key], pd[key], err_msg=f"{name} seed {seed}: {key}")


@pytest.mark.parametrize("sign", [1, -1])
def test_out_of_interval_sample_deletes_the_whole_evaluated_view(sign):
    """reference field.py:31-44 + the DeleteParticle handler: one particle released outside the fields' time interval raises
    OutsideTimeInterval inside the vectorised eval, which flags EVERY particle evaluated in that iteration -- the handler
    then deletes them all; only particles that were not being evaluated (released after this call's endtime) survive."""
    from oracle import parcels_oracle as po

    fs = _fieldset()
    t_out, t_in, t_late = (-3.0, 0.0, 9.0) if sign > 0 else (103.0, 100.0, 91.0)
    x = np.array([100.0, 200.0, 300.0, 400.0])
    t = np.array([t_in, t_out, t_in + sign * 2.0, t_late])
    with pytest.warns(pb.ParticleSetWarning):  # "released outside the FieldSet's executable time domain" (particleset.py:153-160)
        ps = pb.ParticleSet(fs, x=x, y=np.full(4, 100.0), t=t)
    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], dt=sign * 1.0, runtime=10.0)  # endtime = t_out + sign * 10
    assert ps.last_report["n_out_of_time"] == 1
    np.testing.assert_array_equal(ps.particle_id, [3])
    np.testing.assert_array_equal(ps.x, np.float32([400.0]))
    np.testing.assert_array_equal(ps.t, [t_late])
    # ... which is what the oracle's restatement of the reference does
    U = np.ones((2, 1, 11, 21), dtype=np.float32)
    ofs = po.OFieldSet(po.OGrid(np.linspace(0.0, 1000.0, 21), np.linspace(0.0, 500.0, 11), None, mesh="flat"), U, 0 * U, None,
                       time=np.array([0.0, 100.0]))  # fmt: skip
    pd = po.create_particle_data(x, np.full(4, 100.0), np.zeros(4), t)
    po.pset_execute(pd, ofs, [po.AdvectionRK4, po.DeleteOnError], sign * 1.0, runtime=10.0)
    np.testing.assert_array_equal(pd["particle_id"], [3])


@pytest.mark.parametrize("interp", ["linear", "cgrid_velocity"])
def test_first_level_particles_in_a_mixed_batch_are_promoted_like_the_reference(interp):
    """float32 grid, repeated release: half of the set starts exactly on the first time level (tau == 0), the rest later.  The
    reference decides `lenT = 2 if any(tau > 0)` for the whole batch (_xinterpolators.py:130), so the first-level particles'
    first sample is float64 arithmetic -- alone in a batch it would be float32.  Bit-exact against the oracle, both through the
    sampling API (exact batch flags from the pre-pass) and through ParticleSet.execute (the host tells the kernel)."""
    import cases
    from engine_run import make_fieldset, run_engine
    from oracle import parcels_oracle as po
    from oracle_run import oracle_fieldset, run_oracle

    spec = dict(seed=5, kind="smooth", cdtype="f4", ddtype="f4", mesh="flat", nx=20, ny=15, nz=4, nt=4, tstep=600.0, n=4000,
                kernels=["AdvectionRK4_3D"], dt=100.0, segments=[dict(runtime=100.0)], delete=True, margin=0.1, umax=3.0)  # fmt: skip
    if interp != "linear":
        spec["interp"] = interp
    c = cases.build(spec)
    n = len(c["x"])
    c["t"] = np.where(np.arange(n) % 2 == 0, 0.0, 50.0)
    # one evaluation of the whole batch
    fs, ofs = make_fieldset(c), oracle_fieldset(c)
    x32, y32, z32 = (np.asarray(c[k], dtype=np.float32) for k in "xyz")
    got = fs.UVW.eval(c["t"], z32, y32, x32)
    want = po.eval_uvw(ofs, c["t"], z32, y32, x32, None, True)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    alone = fs.UVW.eval(c["t"][::2], z32[::2], y32[::2], x32[::2])  # the first-level particles by themselves: float32 values
    assert np.any(alone[0] != got[0][::2]) and np.all(alone[0] == alone[0].astype(np.float32))
    # the whole run
    ps, err = run_engine(c)
    pd, oerr = run_oracle(c)
    assert not err and oerr is None and len(ps) == len(pd["x"])
    for k in ("particle_id", "state", "t", "ei", "x", "y", "z"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)


@pytest.mark.parametrize("interp", ["freeslip", "partialslip"])
def test_slip_land_test_uses_the_depth_levels_of_the_batch(interp):
    """XFreeslip / XPartialslip look for land on `lenZ` depth levels, and the reference decides `lenZ = 2 if any(zeta > 0)` per
    batch (_xinterpolators.py:401,426-447): a particle exactly on the first depth level is treated differently when another
    particle of the set lies deeper.  Field with 'land' (U = V = 0) at the first depth level only; half of the particles on that
    level, half between the first two: bit-exact against the oracle (the host passes PB_BATCH_TWO_Z)."""
    import cases
    from engine_run import run_engine
    from oracle_run import run_oracle

    c = cases.build(dict(cases.CASES["freeslip_surface"], interp=interp, n=1500, land=False))
    U, V = c["U"].copy(), c["V"].copy()
    half = U.shape[2] // 2
    U[:, 0, :half, :] = 0
    V[:, 0, :half, :] = 0
    c["U"], c["V"] = U, V
    depth = np.asarray(c["depth"], dtype=np.float64)
    c["z"] = np.where(np.arange(len(c["x"])) % 2 == 0, depth[0], 0.5 * (depth[0] + depth[1]))
    ps, err = run_engine(c)
    pd, oerr = run_oracle(c)
    assert not err and oerr is None and len(ps) == len(pd["x"])
    for k in ("particle_id", "state", "t", "ei", "x", "y", "z"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)
    # ... and it is not a no-op: the first-level particles alone (lenZ == 1) end somewhere else
    alone = dict(c)
    for k in ("x", "y", "z", "t"):
        alone[k] = np.asarray(c[k])[::2]
    ps1, _ = run_engine(alone)
    first = ps._data["particle_id"] % 2 == 0
    ids = ps._data["particle_id"][first] // 2
    keep = np.isin(ids, ps1._data["particle_id"])
    assert np.any(ps._data["x"][first][keep] != ps1._data["x"][np.searchsorted(ps1._data["particle_id"], ids[keep])])


@pytest.mark.parametrize("nx, ny, nz", [(1, 7, 5), (9, 1, 5), (1, 1, 4), (6, 7, 1)])
@pytest.mark.parametrize("cdtype", [np.float32, np.float64])
def test_axes_of_length_one(nx, ny, nz, cdtype):
    """A coordinate axis with a single node: the reference's 1-D search returns index 0, coordinate 0 there
    (index_search.py:45-46), every position is 'inside', the field is constant along that axis (reference
    tests/test_advection.py::test_length1dimensions; the oracle agrees with the reference itself bit for bit on such grids)."""
    from oracle import parcels_oracle as po

    rng = np.random.default_rng(3)
    lon = np.linspace(0, 4, nx).astype(cdtype) if nx > 1 else np.array([1.5], dtype=cdtype)
    lat = np.linspace(0, 3, ny).astype(cdtype) if ny > 1 else np.array([0.5], dtype=cdtype)
    depth = np.linspace(0, 10, nz).astype(cdtype) if nz > 1 else np.array([0.0], dtype=cdtype)
    times = np.array([0.0, 100.0, 200.0])
    U, V, W = (rng.uniform(-0.01, 0.01, (3, nz, ny, nx)).astype(np.float32) for _ in range(3))
    n = 50
    x, y, z = rng.uniform(0.2, 3.8, n), rng.uniform(0.2, 2.8, n), rng.uniform(0.5, 9.5, n)
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, depth=depth, time=times, U=U, V=V, W=W, mesh="flat")
    ps = pb.ParticleSet(fs, x=x, y=y, z=z, t=np.zeros(n))
    ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=10.0, runtime=150.0)
    ofs = po.OFieldSet(po.OGrid(lon, lat, depth, mesh="flat"), U, V, W, time=times)
    pd = po.create_particle_data(x, y, z, np.zeros(n))
    po.pset_execute(pd, ofs, [po.AdvectionRK4_3D, po.DeleteOnError], 10.0, runtime=150.0)
    assert len(ps) == len(pd["x"]) > 0
    for k in ("particle_id", "state", "t", "ei", "x", "y", "z"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)


@pytest.mark.parametrize("name", ["curv_flat_2d", "curv_sph_2d"])
def test_lone_particle_in_the_first_cell_column_always_goes_through_the_hash(name):
    """`if np.any(xi)` (reference index_search.py:269): when every hinted xi of a batch is 0 the hint test is skipped and the
    spatial hash answers (float32-rounded cell coordinates).  A set of ONE particle sitting in the first column of a
    curvilinear grid is such a batch at every evaluation, not only at the first of a call."""
    import cases
    from engine_run import run_engine, ulp_diff_f32
    from oracle_run import run_oracle

    c0 = cases.build(cases.CASES[name])
    lon2, lat2 = np.asarray(c0["lon"]), np.asarray(c0["lat"])
    j = lon2.shape[0] // 2
    c = dict(c0, U=c0["U"] * 0.02, V=c0["V"] * 0.02)  # slow flow: the particle stays in column 0 for the whole run
    c["x"], c["y"] = np.array([0.5 * (lon2[j, 0] + lon2[j + 1, 1])]), np.array([0.5 * (lat2[j, 0] + lat2[j + 1, 1])])
    c["z"], c["t"] = np.asarray(c0["z"])[:1], np.asarray(c0["t"])[:1]
    ps, err = run_engine(c)
    pd, oerr = run_oracle(c)
    assert not err and oerr is None and len(ps) == len(pd["x"]) == 1
    assert ps._data["ei"][0, -1] == pd["ei"][0, -1] and ps._data["ei"][0, -1] % (lon2.shape[1] - 1) == 0
    for k in "xy":
        assert ulp_diff_f32(ps._data[k], pd[k]).max() <= (0 if name == "curv_flat_2d" else 8), k
