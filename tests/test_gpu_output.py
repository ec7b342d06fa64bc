"""GPU parity of the output path: device-side ParticleFile.write row selection, ordered
compaction of deleted particles, device-resident output intervals -- against the oracle, the host rule and
the rows the reference itself writes (tests/golden/output_rows.npz).  Bit-exact throughout (flat cases)."""

import os

import numpy as np
import pytest

import parcels_b200 as pb
from engine_run import make_fieldset, ulp_diff_f32
from oracle import parcels_oracle as po
from oracle_run import load_case, oracle_fieldset
from parcels_b200.particle import create_particle_data
from parcels_b200.particlefile import to_write_particles

pytestmark = pytest.mark.gpu


def _upload_random(eng, n, seed, frac_deleted=0.0):
    rng = np.random.default_rng(seed)
    d = create_particle_data(nparticles=n, ngrids=1, initial=dict(
        x=rng.uniform(0, 1, n), y=rng.uniform(0, 1, n), z=rng.uniform(0, 1, n), t=np.round(rng.uniform(-50, 250, n), 1),
        particle_id=rng.permutation(n)))  # fmt: skip
    d["t"][rng.uniform(size=n) < 0.05] = np.nan
    d["t"][rng.uniform(size=n) < 0.02] = np.inf
    d["t"][rng.uniform(size=n) < 0.02] = -np.inf
    d["state"][rng.uniform(size=n) < frac_deleted] = 30
    eng.upload_particles(d, np.ascontiguousarray(d["ei"][:, -1]))
    return d


@pytest.mark.parametrize("n", [1, 255, 256, 257, 100_003, 1_500_000])
def test_device_selection_matches_host_rule(n):
    from parcels_b200.engine import Engine

    eng = Engine(0)
    d = _upload_random(eng, n, seed=n)
    for tout, dt in ((100.0, 10.0), (100.0, -10.0), (104.9, 10.0), (0.0, 600.0), (1e9, 1.0)):
        d["dt"][:] = dt
        rows = to_write_particles(d, tout)
        np.testing.assert_array_equal(rows, po.to_write_particles(d, tout))
        m = eng.output_select(tout, dt)
        assert m == len(rows)
        got = eng.output_gather(m, with_index=True)
        np.testing.assert_array_equal(got["index"], rows)
        for k in ("x", "y", "z", "t", "particle_id"):
            np.testing.assert_array_equal(got[k], d[k][rows])
        part = eng.output_gather(m, columns=("t", "x"))
        assert set(part) == {"t", "x"}
        np.testing.assert_array_equal(part["x"], d["x"][rows])


def test_device_selection_rule_on_reference_vectors(golden_dir):
    from parcels_b200.engine import Engine

    g = np.load(os.path.join(golden_dir, "output_rows.npz"))
    t = g["rule/t"]
    n = len(t)
    eng = Engine(0)
    d = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=np.zeros(n), y=np.zeros(n), z=np.zeros(n), t=t,
                                                                  particle_id=np.arange(n)))  # fmt: skip
    eng.upload_particles(d, np.zeros(n, dtype=np.int32))
    k = 0
    while f"rule/tout{k}" in g:
        # the golden vectors mix dt = +10 and -10 per particle: |dt/2| is what the rule uses
        m = eng.output_select(float(g[f"rule/tout{k}"]), 10.0)
        np.testing.assert_array_equal(eng.output_gather(m, columns=(), with_index=True)["index"], g[f"rule/rows{k}"])
        k += 1


@pytest.mark.parametrize("n,frac", [(1000, 0.3), (70_001, 0.01), (70_001, 0.999), (513, 1.0), (513, 0.0)])
def test_remove_deleted_keeps_order(n, frac):
    from parcels_b200.engine import Engine

    eng = Engine(0)
    d = _upload_random(eng, n, seed=n + int(frac * 1000), frac_deleted=frac)
    keep = d["state"] != 30
    assert eng.remove_deleted() == int(keep.sum())
    out = eng.download_all()
    assert len(out["x"]) == int(keep.sum())
    for k in ("x", "y", "z", "t", "particle_id", "state"):
        np.testing.assert_array_equal(out[k], d[k][keep])


class _Rows:
    """In-memory stand-in for a ParticleFile that uses the same column source as ParticleFile.write."""

    def __init__(self, outputdt):
        self.outputdt, self.rows, self.on_device = outputdt, [], []

    def write(self, pset, t):
        cols, dev = pset._output_columns(float(t), ["t", "z", "y", "x", "particle_id"])
        self.rows.append((float(t), cols))
        self.on_device.append(dev)


@pytest.mark.parametrize("name", ["delayed_partial", "flat_f32c_f64d", "backward"])
def test_resident_intervals_write_the_rows_the_reference_writes(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "output_rows.npz"))
    c = load_case(name)
    outputdt = float(g[f"{name}/outputdt"])
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    rec = _Rows(outputdt)
    ps.execute([getattr(pb, c["kernels"][0]), pb.DeleteParticle], dt=c["dt"], output_file=rec, **c["segments"][0])
    assert rec.on_device[0] is False and all(rec.on_device[1:])  # initial condition from the host, the rest from HBM
    np.testing.assert_array_equal([r[0] for r in rec.rows], g[f"{name}/times"])
    exact = c["mesh"] == "flat"
    # ulps of max(|x|, 1 % of the coordinate range): the longitudes of the spherical case cross zero, where the spacing of
    # the value itself says nothing about the accuracy of the increments that produced it
    floor = {k: 0.01 * float(np.abs(np.asarray(c[k])).max()) or None for k in "xyz"}
    for i, (_, cols) in enumerate(rec.rows):
        np.testing.assert_array_equal(cols["particle_id"], g[f"{name}/{i}/particle_id"])
        np.testing.assert_array_equal(cols["t"], g[f"{name}/{i}/t"])
        for k in "xyz":
            if exact:
                np.testing.assert_array_equal(cols[k], g[f"{name}/{i}/{k}"])
            else:
                assert ulp_diff_f32(cols[k], g[f"{name}/{i}/{k}"], floor=floor[k]).max() <= 2
    # the lazily refreshed host arrays equal the oracle's final state
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    po.pset_execute(pd, oracle_fieldset(c), [getattr(po, c["kernels"][0]), po.DeleteOnError], c["dt"], outputdt=outputdt,
                    **c["segments"][0])  # fmt: skip
    assert ps._host_stale
    assert len(ps) == len(pd["x"])
    for k in ("particle_id", "t", "state", "ei", "dt"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)
    for k in "xyz":
        if exact:
            np.testing.assert_array_equal(ps._data[k], pd[k])
        else:
            assert ulp_diff_f32(ps._data[k], pd[k], floor=floor[k]).max() <= 2


def test_parquet_file_from_resident_set_equals_host_path(tmp_path):
    c = load_case("flat_f32c_f64d")
    tabs = []
    for lazy in (True, False):
        fs = make_fieldset(c)
        # an extra (unwritten) variable forces the host path: same rows must come out
        pclass = pb.Particle if lazy else pb.Particle.add_variable(pb.Variable("tmp", np.float32, to_write=False))
        ps = pb.ParticleSet(fs, pclass=pclass, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
        pf = pb.ParticleFile(tmp_path / f"o{int(lazy)}.parquet", outputdt=20.0)
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=c["dt"], runtime=150.0, output_file=pf)
        assert pf._writer is None  # closed by execute, like `with output_file:`
        assert (pf.device_writes > 0) == lazy
        tabs.append(pb.read_particlefile(pf.path))
    assert list(tabs[0]) == ["t", "z", "y", "x", "particle_id"]
    for k in tabs[0]:
        np.testing.assert_array_equal(tabs[0][k], tabs[1][k])
    assert len(tabs[0]["t"]) > 2000


def test_errors_still_raise_where_the_reference_raises():
    """No error handler: the device-resident path replays from its HBM snapshot and leaves every particle where
    the reference leaves it (kernel.py:239-245)."""
    c = load_case("raise_oob")
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    rec = _Rows(2 * c["dt"])
    with pytest.raises(pb.FieldOutOfBoundError):
        ps.execute([getattr(pb, c["kernels"][0])], dt=c["dt"], output_file=rec, **c["segments"][0])
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    with pytest.raises(po.OracleParticleError):
        po.pset_execute(pd, oracle_fieldset(c), [getattr(po, c["kernels"][0])], c["dt"], outputdt=2 * c["dt"], **c["segments"][0])
    for k in ("particle_id", "t", "state", "ei", "x", "y", "z"):
        np.testing.assert_array_equal(ps._data[k], pd[k], err_msg=k)
