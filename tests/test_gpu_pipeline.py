"""GPU: ``pb_advect_host`` -- Kernel.execute on host arrays cut into pipelined chunks (copy-in, snapshot, kernel, copy-out per
chunk in its own stream) -- leaves exactly what upload + advect + download leave: same particles, same report, same error
replay, with and without the copy-back, on every kind of kernel list the single-launch path serves."""

import numpy as np
import pytest

import parcels_b200 as pb
from engine_run import make_fieldset
from oracle_run import load_case

pytestmark = pytest.mark.gpu

KEYS = ("particle_id", "state", "t", "dt", "ei", "x", "y", "z", "dx", "dy", "dz")


def _run(c, chunks, eager, monkeypatch, handler=True, n_exec=None):
    from parcels_b200.engine import Engine

    monkeypatch.setattr(pb.ParticleSet, "PIPELINE_MIN_PARTICLES", 1)
    calls = []
    inner = Engine.advect_host
    monkeypatch.setattr(Engine, "advect_host", lambda self, *a, **k: (calls.append(k["n_chunks"]), inner(self, *a, **k))[1])
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"], seed=11)
    ps.pipeline_chunks = chunks
    ps.eager_host = eager
    kernels = [getattr(pb, k) for k in c["kernels"]] + ([pb.DeleteParticle] if handler else [])
    err = None
    try:
        for seg in (c["segments"] if n_exec is None else c["segments"][:n_exec]):
            ps.execute(kernels, dt=c["dt"], **seg)
    except Exception as e:  # noqa: BLE001 -- compared between the two paths
        err = type(e).__name__
    assert (len(calls) > 0) == (chunks > 1), "the pipelined entry point is used exactly when asked for"
    return ps, err


def _tile(c, reps):
    """the golden cases are small: repeat the releases so that several 37888-particle chunks exist"""
    c = dict(c)
    for k in ("x", "y", "z", "t"):
        c[k] = np.tile(np.asarray(c[k]), reps)
    return c


@pytest.mark.parametrize("name", ["delayed_partial", "flat_f32c_f64d", "backward", "c2_small", "diffusion", "curv_sph_2d", "cgrid_rect_3d", "freeslip_3d", "raise_time"])
@pytest.mark.parametrize("eager", [False, True])
def test_pipelined_equals_single_launch(name, eager, monkeypatch):
    c = load_case(name)
    c = _tile(c, max(1, 120_000 // len(c["x"])))
    ref, rerr = _run(c, 0, eager, monkeypatch)
    got, gerr = _run(c, 3, eager, monkeypatch)
    assert rerr == gerr
    assert got.last_report["particle_steps"] == ref.last_report["particle_steps"] > 0
    for k in ("n_deleted", "n_error", "max_state", "cache_refills"):
        assert got.last_report[k] == ref.last_report[k], k
    assert len(got) == len(ref)
    for k in KEYS:
        np.testing.assert_array_equal(got._data[k], ref._data[k], err_msg=k)


def test_pipelined_error_replay_leaves_the_reference_state(monkeypatch):
    """no handler: the first error stops the whole set at the end of that iteration (kernel.py:239-245) -- replayed from the
    per-chunk snapshot the pipelined call took, also when the (now stale) result had already been copied back"""
    c = _tile(load_case("raise_oob"), 400)
    for eager in (False, True):
        ref, rerr = _run(c, 0, eager, monkeypatch, handler=False)
        got, gerr = _run(c, 4, eager, monkeypatch, handler=False)
        assert rerr == gerr == "FieldOutOfBoundError"
        for k in KEYS:
            np.testing.assert_array_equal(got._data[k], ref._data[k], err_msg=k)


def test_chunking_edge_sizes(monkeypatch):
    """n below one chunk, exactly one chunk, one more than a chunk, more chunks than streams"""
    c0 = load_case("flat_f32c_f64d")
    for n in (1, 37_888, 37_889, 37_888 * 5 + 17):
        c = dict(c0)
        idx = np.arange(n) % len(c0["x"])
        for k in ("x", "y", "z", "t"):
            c[k] = np.asarray(c0[k])[idx]
        ref, _ = _run(c, 0, True, monkeypatch, n_exec=1)
        for chunks in (2, 7):
            got, _ = _run(c, chunks, True, monkeypatch, n_exec=1)
            assert len(got) == len(ref)
            for k in KEYS:
                np.testing.assert_array_equal(got._data[k], ref._data[k], err_msg=f"{k} n={n} chunks={chunks}")


def test_dt_column_after_execute_with_the_deferred_fill():
    """`particles.dt = dt` (reference _core/particleset.py:414, kernel.py:225-226): for fused plans on large sets the host column is
    filled by a helper thread under the first pipelined launch -- whatever path execute() takes, the column must hold dt afterwards
    (pipelined eager / lazy, a zero-length run, a backward run) and a user kernel must see it filled up front."""
    import bench

    f = bench.c2_field(nx=60, ny=40, nz=6, nt=3)
    fs = pb.FieldSet.from_arrays(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["times"], U=f["U"], V=f["V"], W=f["W"], mesh="spherical")
    n = 300_000
    p = bench.c2_particles(f, n, 1)
    for eager in (True, False):
        ps = pb.ParticleSet(fs, x=p["x"], y=p["y"], z=p["z"], t=p["t"], seed=1)
        ps.eager_host = eager
        ps._data["dt"][:] = 123.0
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=600.0, runtime=600.0)
        assert np.all(ps._data["dt"] == 600.0) and "_dt_pending" not in ps.__dict__
        ps._data["dt"][:] = 5.0
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=300.0, runtime=0.0)
        assert np.all(ps._data["dt"] == 300.0)
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=-300.0, runtime=300.0)
        assert np.all(ps._data["dt"] == -300.0)
    ps = pb.ParticleSet(fs, x=p["x"][:1000], y=p["y"][:1000], z=p["z"][:1000], t=p["t"][:1000])
    seen = []

    def K(particles, fieldset):
        seen.append(np.unique(np.asarray(particles.dt)).tolist())

    ps.execute([pb.AdvectionRK4_3D, K], dt=600.0, runtime=600.0)
    assert seen and seen[0] == [600.0]
