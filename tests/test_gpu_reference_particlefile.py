"""The reference's own ``tests/test_particlefile.py``, transcribed (same names, same assertions) for the part that goes through
``ParticleSet.execute(..., output_file=ParticleFile(...))`` and ``ParticleFile.write``.  The `t` column is read as the stored
float64 seconds (the reference's reader decodes it to timestamps with polars, which is not installed here; the CF attributes
that decoding needs are checked by tests/test_output_cpu.py).  Left out: tests marked skip / xfail in the reference and the one whose body is `...`."""

from contextlib import nullcontext as does_not_raise
from datetime import datetime, timedelta

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import parcels_b200 as pb
from parcels_b200 import AdvectionRK4, Particle, ParticleFile, ParticleSet, ParticleSetWarning, Variable
from test_gpu_reference_execute import TIME, DoNothing, fieldset  # noqa: F401 -- the fixture

pytestmark = pytest.mark.gpu


@pytest.fixture
def tmp_parquet(tmp_path):
    return tmp_path / "tmp.parquet"


def test_metadata(fieldset, tmp_parquet):
    pset = ParticleSet(fieldset, pclass=Particle, x=0, y=0)
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset.execute(DoNothing, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    tab = pq.read_table(tmp_parquet)
    assert tab.schema.metadata[b"parcels_kernels"].decode().lower() == "DoNothing".lower()


@pytest.mark.parametrize("compression", ["zstd", "gzip", "snappy", "brotli", None])
def test_compression(fieldset, tmp_parquet, compression):
    pset = ParticleSet(fieldset, pclass=Particle, x=0, y=0)
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"), compression=compression)
    pset.execute(DoNothing, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    tab = pq.ParquetFile(tmp_parquet)
    for i in range(tab.num_row_groups):
        row_group = tab.metadata.row_group(i)
        for j in range(row_group.num_columns):
            col = row_group.column(j)
            assert col.compression.lower() == compression or (compression is None and col.compression.lower() == "uncompressed")


def test_write_fieldset_without_time(tmp_parquet):
    z = np.zeros((1, 1, 20, 30))  # a dataset without a time dimension
    fieldset = pb.FieldSet.from_arrays(lon=np.linspace(0, 1e5, 30), lat=np.linspace(0, 5e4, 20), U=z, V=z, mesh="flat")
    pset = ParticleSet(fieldset, pclass=Particle, x=0, y=0)
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset.execute(DoNothing, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    table = pq.read_table(tmp_parquet)
    assert table.schema.field("t").metadata[b"units"] == b"seconds"
    assert b"calendar" not in table.schema.field("t").metadata
    assert table["t"].to_numpy()[1] == 1.0


def test_pfile_array_remove_particles(fieldset, tmp_parquet):
    """If a particle from the middle of a particleset is removed, that writing doesn't crash"""
    npart = 10
    pset = ParticleSet(fieldset, pclass=Particle, x=np.linspace(0, 1, npart), y=0.5 * np.ones(npart), t=TIME[0])
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset._data["t"][:] = 0
    pfile.write(pset, t=TIME[0])
    pset.remove_indices(3)
    new_time = 86400  # s in a day
    pset._data["t"][:] = new_time
    pfile.write(pset, new_time)
    pfile.close()


def test_pfile_array_remove_all_particles(fieldset, tmp_parquet):
    npart = 10
    pset = ParticleSet(fieldset, pclass=Particle, x=np.linspace(0, 1, npart), y=0.5 * np.ones(npart), t=TIME[0])
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pfile.write(pset, t=0)
    for _ in range(npart):
        pset.remove_indices(-1)
    pfile.write(pset, TIME[0] + np.timedelta64(1, "D"))
    pfile.write(pset, TIME[0] + np.timedelta64(2, "D"))
    pfile.close()
    df = pd.read_parquet(tmp_parquet)
    assert df["particle_id"].nunique() == npart


def test_write_dtypes_pfile(fieldset, tmp_parquet):
    dtypes = [np.float32, np.float64, np.int32, np.uint32, np.int64, np.uint64, np.bool_, np.int8, np.uint8, np.int16, np.uint16]
    extra_vars = [Variable(f"v_{d.__name__}", dtype=d, initial=0.0) for d in dtypes]
    MyParticle = Particle.add_variable(extra_vars)
    pset = ParticleSet(fieldset, pclass=MyParticle, x=0, y=0, t=TIME[0])
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pfile.write(pset, t=TIME[0])
    pfile.close()
    tab = pq.read_table(tmp_parquet)
    for d in dtypes:
        assert tab[f"v_{d.__name__}"].type == pa.from_numpy_dtype(d)


def test_file_warnings(fieldset, tmp_parquet):
    pset = ParticleSet(fieldset, x=[0, 0], y=[0, 0], t=[np.timedelta64(0, "s"), np.timedelta64(1, "s")])
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(2, "s"))
    with pytest.warns(ParticleSetWarning, match="Some of the particles have a start time difference.*"):
        pset.execute(AdvectionRK4, runtime=3, dt=1, output_file=pfile)


@pytest.mark.parametrize("outputdt, expectation", [(np.timedelta64(5, "s"), does_not_raise()), (timedelta(seconds=2), does_not_raise()),
                                                   (5.0, does_not_raise()), (np.datetime64("2001-01-02T00:00:00"), pytest.raises(ValueError)),
                                                   (datetime(2000, 1, 2, 0, 0, 0), pytest.raises(ValueError)),
                                                   (-np.timedelta64(5, "s"), pytest.raises(ValueError))])  # fmt: skip
def test_outputdt_types(outputdt, expectation, tmp_parquet):
    with expectation:
        pfile = ParticleFile(tmp_parquet, outputdt=outputdt)
        want = outputdt.total_seconds() if isinstance(outputdt, timedelta) else (outputdt / np.timedelta64(1, "s") if isinstance(outputdt, np.timedelta64) else outputdt)
        assert pfile.outputdt == want


def test_write_timebackward(fieldset, tmp_parquet):
    release_time = TIME[0] + np.array([np.timedelta64(i + 1, "s") for i in range(3)])
    pset = ParticleSet(fieldset, y=[0, 1, 2], x=[0, 0, 0], t=release_time)
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset.execute(DoNothing, runtime=np.timedelta64(3, "s"), dt=-np.timedelta64(1, "s"), output_file=pfile)
    df = pd.read_parquet(tmp_parquet)
    assert df["particle_id"].dtype == "int64"
    dt_per_particle = df.groupby("particle_id")["t"].diff().dropna()
    assert (dt_per_particle < 0).all()


@pytest.mark.parametrize("outputdt", [np.timedelta64(1, "s"), np.timedelta64(2, "s"), np.timedelta64(3, "s")])
def test_time_is_age(fieldset, tmp_parquet, outputdt):
    npart = 10
    AgeParticle = Particle.add_variable(Variable("age", initial=0.0))

    def IncreaseAge(particles, fieldset):
        particles.age += particles.dt

    time = TIME[0] + np.arange(npart) * np.timedelta64(1, "s")
    pset = ParticleSet(fieldset, pclass=AgeParticle, x=npart * [0], y=npart * [0], t=time)
    ofile = ParticleFile(tmp_parquet, outputdt=outputdt)
    if outputdt > np.timedelta64(1, "s"):
        warning_ctx = pytest.warns(ParticleSetWarning, match="Some of the particles have a start time difference.*")
    else:
        warning_ctx = does_not_raise()
    with warning_ctx:
        pset.execute(IncreaseAge, runtime=np.timedelta64(npart * 2, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    df = pd.read_parquet(tmp_parquet)
    for i, (_, df_traj) in enumerate(df.groupby("particle_id", sort=True)):
        assert (df_traj["age"] == df_traj["t"] - float(i)).all()  # released i seconds after the start of the time axis


@pytest.mark.parametrize("npart", [1, 10])
def test_sampling_initial_value(fieldset, npart, tmp_parquet):
    SampleParticle = Particle.add_variable(Variable("sample", initial=np.nan))

    def SampleKernel(particles, fieldset):
        particles.sample, _ = fieldset.UV[particles]

    x, y = np.zeros(npart), np.zeros(npart)
    t = np.zeros(npart, dtype="timedelta64[s]")
    pset = ParticleSet(fieldset, pclass=SampleParticle, x=x, y=y, t=t)
    pset.sample, _ = fieldset.UV[pset]  # Sample initial value
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset.execute(SampleKernel, runtime=np.timedelta64(2, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    df = pd.read_parquet(tmp_parquet)
    assert np.isfinite(df["sample"]).all()


def test_reset_dt(fieldset, tmp_parquet):
    # p.dt gets reset when a write time is not a multiple of dt: steps of [20, 20, 10, 20, 20, 10] s -> 6 kernel executions
    dt = np.timedelta64(20, "s")

    def Update_lon(particles, fieldset):
        particles.dx += 0.1

    pset = ParticleSet(fieldset, pclass=Particle, x=[0], y=[0])
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(50, "s"))
    pset.execute(Update_lon, runtime=5 * dt, dt=dt, output_file=ofile)
    assert np.allclose(pset.x, 0.6)


@pytest.mark.parametrize("dt", [100, 200])
def test_subsecond_outputdt(fieldset, dt, tmp_parquet):
    def Update_lon(particles, fieldset):
        particles.dx += dt / 1000.0  # Move at a rate of 1 unit per second

    pset = ParticleSet(fieldset, x=[0], y=[0])
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(dt, "ms"))
    pset.execute(Update_lon, runtime=np.timedelta64(1, "s"), dt=np.timedelta64(dt, "ms"), output_file=ofile)
    df = pd.read_parquet(tmp_parquet)
    np.testing.assert_allclose(df["x"], np.arange(0, 1 + 1e-6, dt / 1000.0), atol=1e-6)
    np.testing.assert_allclose((df["t"] - df["t"].min()) * 1000, np.arange(0, 1001, dt), atol=1)


def test_correct_misaligned_outputdt_dt(fieldset, tmp_parquet):
    """outputdt does not need to be a multiple of dt"""

    def Update_lon(particles, fieldset):
        particles.x += particles.dt

    pset = ParticleSet(fieldset, pclass=Particle, x=[0], y=[0])
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(3, "s"))
    pset.execute(Update_lon, runtime=np.timedelta64(11, "s"), dt=np.timedelta64(2, "s"), output_file=ofile)
    df = pd.read_parquet(tmp_parquet)
    assert np.allclose(df["x"].values, [0, 3, 6, 9])
    assert np.allclose(df["t"] - df["t"].min(), [0, 3, 6, 9])


def _setup_pset_execute(fieldset, outputdt, tmp_path, **execute_kwargs):
    npart = 10
    g = fieldset.U.grid
    pset = ParticleSet(fieldset, pclass=Particle, x=np.full(npart, g.lon.mean()), y=np.full(npart, g.lat.mean()))
    name = tmp_path / "tmp.parquet"
    pset.execute(DoNothing, output_file=ParticleFile(name, outputdt=outputdt), **execute_kwargs)
    return pd.read_parquet(name)


def test_pset_execute_outputdt_forwards(fieldset, tmp_path):
    outputdt, runtime, dt = timedelta(hours=1), timedelta(hours=5), timedelta(minutes=5)
    df = _setup_pset_execute(fieldset, outputdt, tmp_path, runtime=runtime, dt=dt)
    np.testing.assert_equal(np.diff(df[df["particle_id"] == 0]["t"]), outputdt.seconds)


def test_pset_execute_output_time_forwards(fieldset, tmp_path):
    outputdt, runtime, dt = np.timedelta64(1, "h"), np.timedelta64(5, "h"), np.timedelta64(5, "m")
    df = _setup_pset_execute(fieldset, outputdt, tmp_path, runtime=runtime, dt=dt)
    assert df["t"].min() == 0.0  # seconds since fieldset.time_interval.left
    assert df["t"].max() - df["t"].min() == runtime / np.timedelta64(1, "s")


def test_pset_execute_outputdt_backwards(fieldset, tmp_path):
    outputdt, runtime, dt = timedelta(hours=1), timedelta(days=2), -timedelta(minutes=5)
    df = _setup_pset_execute(fieldset, outputdt, tmp_path, runtime=runtime, dt=dt)
    np.testing.assert_equal(np.diff(df[df["particle_id"] == 0]["t"]), -outputdt.seconds)


def test_particlefile_init(tmp_parquet):
    ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))


def test_particlefile_init_existing_path_modes(fieldset, tmp_parquet):
    pset = ParticleSet(fieldset, pclass=Particle, x=0, y=0)
    first_file = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    pset.execute(DoNothing, runtime=np.timedelta64(10, "s"), dt=np.timedelta64(1, "s"), output_file=first_file)
    df_first = pd.read_parquet(tmp_parquet)
    with pytest.raises(ValueError, match="already exists"):
        ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    overwrite_file = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"), mode="w")
    pset.execute(DoNothing, runtime=np.timedelta64(10, "s"), dt=np.timedelta64(1, "s"), output_file=overwrite_file)
    df_overwrite = pd.read_parquet(tmp_parquet)
    assert len(df_first) == len(df_overwrite)


def test_particlefile_init_existing_path_no_mode(tmp_parquet):
    tmp_parquet.touch()
    with pytest.raises(ValueError, match="already exists"):
        ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))


def test_particlefile_init_nonexistent_parent(tmp_path):
    path = tmp_path / "nonexistent_dir" / "file.parquet"
    with pytest.raises(ValueError, match="does not exist"):
        ParticleFile(path, outputdt=np.timedelta64(1, "s"))


def test_particlefile_init_invalid_mode(tmp_parquet):
    with pytest.raises(ValueError, match="Invalid mode value"):
        ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"), mode="something-else")


@pytest.mark.parametrize("name", ["path", "outputdt"])
def test_particlefile_readonly_attrs(tmp_parquet, name):
    pfile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    with pytest.raises(AttributeError, match="property .* of 'ParticleFile' object has no setter"):
        setattr(pfile, name, "something")


def test_particlefile_init_invalid(tmp_path):
    path = tmp_path / "file.not-parquet"
    with pytest.raises(ValueError, match="file extension must be '.parquet'"):
        ParticleFile(path, outputdt=np.timedelta64(1, "s"))


def test_particlefile_readable_after_kernel_error(fieldset, tmp_parquet):
    """Parquet output file must be readable even if the kernel raises an error mid-execution (reference GH-2713)."""
    from parcels_b200 import StatusCode

    def ErrorKernel(particles, fieldset):
        particles.state = StatusCode.Error

    pset = ParticleSet(fieldset, pclass=Particle, x=0, y=0)
    ofile = ParticleFile(tmp_parquet, outputdt=np.timedelta64(1, "s"))
    with pytest.raises(RuntimeError, match="General error occurred at"):
        pset.execute(ErrorKernel, runtime=np.timedelta64(10, "s"), dt=np.timedelta64(1, "s"), output_file=ofile)
    df = pd.read_parquet(tmp_parquet)
    assert len(df) >= 1  # at least the initial condition was written


@pytest.mark.parametrize("particle", [
    Particle,
    pb.ParticleClass(variables=[
        Variable("lon", dtype=np.float32, attrs={"standard_name": "longitude", "units": "degrees_east", "axis": "X"}),
        Variable("lat", dtype=np.float32, attrs={"standard_name": "latitude", "units": "degrees_north", "axis": "Y"}),
        Variable("z", dtype=np.float32, attrs={"standard_name": "vertical coordinate", "units": "m", "positive": "down"}),
    ]),
])  # fmt: skip
def test_particle_schema(particle, fieldset, tmp_parquet):
    """the schema of the written table: one field per writable Variable, its attrs as field metadata, CF time attributes on `t`
    (the reference calls its get_schema(particle, {}, TimeInterval(...)); here: the schema ParticleFile builds for a fieldset
    with a datetime time axis)"""
    s = ParticleFile(tmp_parquet, outputdt=1.0)._schema(particle, fieldset)
    written_variables = [v for v in particle.variables if v.to_write]
    assert len(s.names) == len(written_variables)
    for variable, pyarrow_field in zip(written_variables, s, strict=False):
        assert variable.name == pyarrow_field.name
        if variable.name != "t":
            assert variable.attrs == {k.decode(): v.decode() for k, v in pyarrow_field.metadata.items()}
        else:
            assert b"units" in pyarrow_field.metadata
            assert b"calendar" in pyarrow_field.metadata
        assert pa.from_numpy_dtype(variable.dtype) == pyarrow_field.type
