"""Regression tests of the round-1 advisor findings: ParticleSets that share a FieldSet (one engine, one resident SoA) keep
independent data; the adapter knows the extra variables of a shared SoA; mixed NaN / finite release times start like the
reference; a stale snapshot is refused."""

import numpy as np
import pytest

import parcels_b200 as pb

pytestmark = pytest.mark.gpu


def _fieldset():
    lon = np.linspace(0.0, 100.0, 21)
    lat = np.linspace(0.0, 100.0, 21)
    times = np.array([0.0, 100.0])
    U = np.zeros((2, 1, 21, 21), dtype=np.float32)
    V = np.ones((2, 1, 21, 21), dtype=np.float32)  # 1 m/s northward
    return pb.FieldSet.from_arrays(lon=lon, lat=lat, depth=None, time=times, U=U, V=V, W=None, mesh="flat")


def test_two_particlesets_on_one_fieldset_keep_independent_data():
    """reference _core/particleset.py: every ParticleSet owns its `_data`.  After a lazy (device-resident) execute() the only copy
    of set `a` lives in the FieldSet's engine -- executing set `b` must not replace it."""
    fs = _fieldset()
    a = pb.ParticleSet(fs, x=[10.0, 20.0], y=[0.0, 0.0], t=[0.0, 0.0])
    b = pb.ParticleSet(fs, x=[30.0, 40.0], y=[2.0, 2.0], t=[0.0, 0.0])
    a.execute([pb.AdvectionRK4], dt=1.0, runtime=10.0)
    b.execute([pb.AdvectionRK4], dt=1.0, runtime=5.0)
    np.testing.assert_array_equal(a.x, np.float32([10.0, 20.0]))
    np.testing.assert_array_equal(a.y, np.float32([10.0, 10.0]))
    np.testing.assert_array_equal(a.t, [10.0, 10.0])
    np.testing.assert_array_equal(b.y, np.float32([7.0, 7.0]))
    np.testing.assert_array_equal(b.t, [5.0, 5.0])
    # and back: `a` continues from its own state
    a.execute([pb.AdvectionRK4], dt=1.0, runtime=1.0)
    np.testing.assert_array_equal(a.y, np.float32([11.0, 11.0]))
    np.testing.assert_array_equal(b.y, np.float32([7.0, 7.0]))


def test_two_particlesets_of_different_sizes():
    fs = _fieldset()
    c = pb.ParticleSet(fs, x=[10.0, 20.0], y=[0.0, 0.0], t=[0.0, 0.0], particle_ids=[100, 101])
    d = pb.ParticleSet(fs, x=[30.0, 40.0, 50.0], y=[1.0, 1.0, 1.0], t=[0.0, 0.0, 0.0], particle_ids=[7, 8, 9])
    c.execute([pb.AdvectionRK4], dt=1.0, runtime=4.0)
    d.execute([pb.AdvectionRK4], dt=1.0, runtime=4.0)
    assert len(c) == 2 and len(c.x) == 2
    np.testing.assert_array_equal(c.particle_id, [100, 101])
    np.testing.assert_array_equal(c.y, np.float32([4.0, 4.0]))
    np.testing.assert_array_equal(d.particle_id, [7, 8, 9])
    np.testing.assert_array_equal(d.y, np.float32([5.0, 5.0, 5.0]))


def test_interleaved_output_intervals_of_two_sets():
    """`a` is read (output rows selected on the device) while `b` has run in between."""
    fs = _fieldset()
    a = pb.ParticleSet(fs, x=[10.0], y=[0.0], t=[0.0])
    b = pb.ParticleSet(fs, x=[30.0], y=[50.0], t=[0.0])
    a.execute([pb.AdvectionRK4, pb.DeleteParticle], dt=1.0, runtime=3.0)
    b.execute([pb.AdvectionRK4, pb.DeleteParticle], dt=1.0, runtime=60.0)  # leaves the domain at y = 100: deleted
    assert len(b) == 0
    assert len(a) == 1 and float(a.y[0]) == 3.0


def test_adapter_shares_extra_variables_through_deletions():
    """adapter.pset_from_parcels shares the reference's `_data` dict, which may hold extra Variables: a deletion must shrink EVERY
    column (reference Kernel.remove_deleted -> remove_indices over all keys, _core/kernel.py:98-106)."""
    from types import SimpleNamespace

    from parcels_b200.adapter import pset_from_parcels
    from parcels_b200.particle import create_particle_data

    fs = _fieldset()
    n = 4
    d = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=np.array([10.0, 20.0, 30.0, 40.0]), y=np.array([1.0, 99.5, 2.0, 3.0]),
                                                                  z=np.zeros(n), t=np.zeros(n), particle_id=np.arange(n)))  # fmt: skip
    d["age"] = np.array([1.0, 2.0, 3.0, 4.0], dtype=np.float32)
    ps = pset_from_parcels(SimpleNamespace(_data=d), fs)
    assert [v.name for v in ps._pclass.extra] == ["age"]
    ps.execute([pb.AdvectionRK4, pb.DeleteParticle], dt=1.0, runtime=5.0)  # the particle at y = 99.5 leaves the domain
    assert d["x"].shape == (3,) and d["particle_id"].tolist() == [0, 2, 3]
    assert d["age"].shape == (3,) and d["age"].tolist() == [1.0, 3.0, 4.0]
    assert d["ei"].shape == (3, 1)


def test_mixed_nan_and_finite_release_times_start_at_the_fieldset_start():
    """reference _core/particleset.py:541-544,575-585,413-414: `release_times.min()` is NaN as soon as one time is unset, the
    start time is then the fieldset's start and every particle's t is overwritten with it."""
    fs = _fieldset()
    ps = pb.ParticleSet(fs, x=[10.0, 20.0, 30.0], y=[0.0, 0.0, 0.0], t=[np.nan, 5.0, 7.0])
    ps.execute([pb.AdvectionRK4], dt=1.0, runtime=3.0)
    np.testing.assert_array_equal(ps.t, [3.0, 3.0, 3.0])
    np.testing.assert_array_equal(ps.y, np.float32([3.0, 3.0, 3.0]))
    # backward in time: the start is the END of the interval
    ps = pb.ParticleSet(fs, x=[10.0, 20.0], y=[50.0, 50.0], t=[np.nan, 5.0])
    ps.execute([pb.AdvectionRK4], dt=-1.0, runtime=3.0)
    np.testing.assert_array_equal(ps.t, [97.0, 97.0])


def test_restore_refuses_a_snapshot_of_another_set_size():
    from parcels_b200._lib import EngineError
    from parcels_b200.particle import create_particle_data

    fs = _fieldset()
    eng = fs.engine(0)
    d = create_particle_data(nparticles=4, ngrids=1, initial=dict(x=np.full(4, 10.0), y=np.full(4, 1.0), z=np.zeros(4), t=np.zeros(4),
                                                                  particle_id=np.arange(4)))  # fmt: skip
    eng.upload_particles(d, np.zeros(4, dtype=np.int32))
    eng.snapshot()
    eng.restore()
    d3 = {k: v[:3].copy() for k, v in d.items()}
    eng.upload_particles(d3, np.zeros(3, dtype=np.int32))  # another set: the snapshot of the old one is void
    with pytest.raises(EngineError):
        eng.restore()


def test_snapshot_survives_compaction_and_restores_count_and_ids():
    """pb_particles_snapshot holds the whole set (ids included): after pb_particles_remove_deleted changed the count, restore
    brings back the pre-compaction set -- what a domain-decomposed bench pass needs (migration changes the count every pass)."""
    from parcels_b200.particle import create_particle_data
    from parcels_b200.statuscodes import StatusCode

    fs = _fieldset()
    eng = fs.engine(0)
    d = create_particle_data(nparticles=5, ngrids=1, initial=dict(x=np.linspace(10, 50, 5), y=np.full(5, 1.0), z=np.zeros(5), t=np.zeros(5),
                                                                  particle_id=np.array([11, 12, 13, 14, 15])))  # fmt: skip
    d["state"][[1, 3]] = StatusCode.Delete
    eng.upload_particles(d, np.zeros(5, dtype=np.int32))
    eng.snapshot()
    assert eng.remove_deleted() == 3
    assert eng.download_all()["particle_id"].tolist() == [11, 13, 15]
    eng.restore()
    back = eng.download_all()
    assert eng.particle_count() == 5 and back["particle_id"].tolist() == [11, 12, 13, 14, 15]
    np.testing.assert_array_equal(back["x"], d["x"])
    np.testing.assert_array_equal(back["state"], d["state"])
