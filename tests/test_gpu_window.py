"""GPU: time-slab streaming (only `time_window` levels of U, V, W resident, next level prefetched while the
kernel runs) gives bit-identical trajectories to the fully resident field -- forward, backward, across several
execute() segments, with delayed releases."""

import numpy as np
import pytest

import bench
import parcels_b200 as pb

pytestmark = pytest.mark.gpu


def _field(nt=7):
    f = bench.c2_field(nx=90, ny=45, nz=10, nt=nt)
    f["times"] = np.arange(nt) * 3600.0
    f["U"] *= np.float32(20.0)
    f["V"] *= np.float32(20.0)
    return f


def _run(f, window, dt, segments, t0, n=4000, seed=3):
    rng = np.random.default_rng(seed)
    x, y, z = rng.uniform(-170, 170, n), rng.uniform(-70, 70, n), rng.uniform(5, 5000, n)
    fs = pb.FieldSet.from_arrays(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["times"], U=f["U"], V=f["V"], W=f["W"],
                                 mesh="spherical", time_window=window)  # fmt: skip
    ps = pb.ParticleSet(fs, x=x, y=y, z=z, t=t0(n, rng))
    for seg in segments:
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=dt, **seg)
    return ps


@pytest.mark.parametrize("window", [2, 3])
@pytest.mark.parametrize("mode", ["forward", "backward", "delayed"])
def test_windowed_equals_resident(window, mode):
    f = _field()
    tend = float(f["times"][-1])
    # window == 2 can only serve steps that never straddle a time level (a straddling step samples 3 levels):
    # those cases keep every step aligned with the hourly levels; window == 3 also gets unaligned segments
    if mode == "forward" and window == 2:
        dt, segs, t0 = 600.0, [dict(runtime=2 * 3600.0), dict(runtime=3 * 3600.0 + 300.0)], lambda n, r: np.zeros(n)
    elif mode == "forward":
        dt, segs, t0 = 600.0, [dict(runtime=0.4 * tend), dict(runtime=0.55 * tend)], lambda n, r: np.zeros(n)
    elif mode == "backward":
        dt, segs, t0 = -600.0, [dict(runtime=0.9 * tend)], lambda n, r: np.full(n, tend)
    else:
        q = 900 if window == 2 else 300
        dt, segs, t0 = 900.0, [dict(endtime=0.8 * tend)], lambda n, r: np.round(r.uniform(0, 0.5 * tend, n) / q) * q
    ref = _run(f, None, dt, segs, t0)
    win = _run(f, window, dt, segs, t0)
    assert win.last_report["particle_steps"] > 0
    for k in ("particle_id", "state", "t", "ei", "x", "y", "z"):
        np.testing.assert_array_equal(win._data[k], ref._data[k], err_msg=k)


def test_window_too_small_for_dt_is_an_error():
    f = _field()
    with pytest.raises(RuntimeError, match="widen time_window"):
        _run(f, 2, 5400.0, [dict(runtime=3 * 5400.0)], lambda n, r: np.zeros(n))  # one step spans 1.5 level intervals


def test_windowed_diffusion_draws_fresh_increments_after_every_window_slide():
    """Fused DiffusionUniformKh on a time-windowed field: a launch resumed after a window slide restarts its iteration count,
    so it must not draw the Wiener increments of the previous launch again (Var = 2 K t; repeating the increments of each
    1-hour window over 4 windows would double the standard deviation)."""
    nt, n, K = 5, 4000, 100.0
    lon, lat = np.linspace(-1e5, 1e5, 9), np.linspace(-1e5, 1e5, 7)
    Z = np.zeros((nt, 1, 7, 9), dtype=np.float32)
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, time=np.arange(nt) * 3600.0, U=Z, V=Z, mesh="flat", time_window=2)
    fs.add_constant_field("Kh_zonal", K, mesh="flat")
    fs.add_constant_field("Kh_meridional", K, mesh="flat")
    ps = pb.ParticleSet(fs, x=np.zeros(n), y=np.zeros(n), t=np.zeros(n), seed=5)
    runtime = 4 * 3600.0
    ps.execute([pb.AdvectionRK4, pb.DiffusionUniformKh, pb.DeleteParticle], dt=600.0, runtime=runtime)
    assert len(ps) == n and np.all(ps.t == runtime)
    sigma = np.sqrt(2 * K * runtime)
    assert abs(np.std(ps.x) / sigma - 1) < 0.05 and abs(np.std(ps.y) / sigma - 1) < 0.05, (np.std(ps.x), np.std(ps.y), sigma)
