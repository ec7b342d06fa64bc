"""Differential fuzzing of round 2's additions, host-compiled kernel sources (oracle/hostsim) against the oracle:
  curv      AdvectionRK4 / RK4_3D on random CURVILINEAR meshes (flat / spherical, f32 / f64 nodes, 2-D / 3-D) with CGrid_Velocity and
            with XLinear_Velocity, particles released near the rim, several execute() segments;
  curv_s    scalar Field.eval (XLinear / XNearest / CGrid_Tracer) on such meshes, fresh set + hinted second evaluation;
  advdiff_c AdvectionDiffusionM1 / EM with CGrid_Velocity on rectilinear C-grids (the oracle fed the device's own increments);
  mig       in-kernel migration: 2-5 slab engines of one process linked by address, random inbox capacity (overflow), against the
            undecomposed run, bit for bit;
  mig_w     the same under time-slab streaming, through distributed.execute_decomposed itself on THREAD ranks (scripts/thread_ranks.py):
            random window / level count / direction / staggered releases, against the undecomposed resident run;
  multigrid a scalar and a vector field on further XGrids sampled through their own engines.
Run:  PB_LIB=oracle/_build/hostsim/libparcels_b200_hostsim.so PB_HOSTSIM_TEST=1 python scripts/fuzz_hostsim_r2.py [n] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import warnings
import numpy as np
import cases
import parcels_b200 as pb
from engine_run import make_fieldset, ulp_diff_f32
from oracle import parcels_oracle as po
from oracle_run import oracle_fieldset
from parcels_b200 import distributed as D
from parcels_b200.fieldset import XGrid
from parcels_b200.particle import create_particle_data
from philox_ref import device_normals

warnings.simplefilter("ignore")
CURV_ULP = 8  # the curvilinear tolerance of tests/test_gpu_parity.py


def curv_case(rng, interp=None):
    mesh = str(rng.choice(["flat", "spherical"]))
    three = bool(rng.random() < 0.4)
    interp = interp or str(rng.choice(["cgrid_velocity", "linear"]))
    kern = "AdvectionRK4_3D" if three else "AdvectionRK4"
    spec = dict(seed=int(rng.integers(1, 10**6)), kind="curv", interp=interp, cdtype=str(rng.choice(["f4", "f8"])), mesh=mesh,
                nx=int(rng.integers(8, 30)), ny=int(rng.integers(7, 24)), nz=int(rng.integers(3, 7)) if three else int(rng.choice([1, 4])),
                nt=int(rng.choice([2, 3])), tstep=float(rng.choice([600.0, 3600.0])), n=int(rng.integers(1, 150)), kernels=[kern],
                dt=60.0, segments=[dict(runtime=60.0)], delete=True, umax=1.0)  # fmt: skip
    c = cases.build(spec)
    # speeds that cross a few cells per run, whatever the mesh units
    cell = float(np.abs(np.diff(np.asarray(c["lon"], dtype=np.float64), axis=1)).mean())
    nstep = int(rng.integers(3, 12))
    dt = float(rng.choice([300.0, 900.0])) * (1 if rng.random() < 0.8 else -1)
    per_step = cell * float(rng.choice([0.05, 0.2, 0.6]))
    scale = per_step / abs(dt) * (111e3 if mesh == "spherical" else 1.0)  # (degrees per step -> m/s on a spherical mesh)
    for k in "UV":
        c[k] = (c[k] * np.float32(scale)).astype(np.float32)
    if not three:
        c["W"] = None
    if rng.random() < 0.5:  # some particles next to / outside the rim: out-of-bounds + hash misses
        m = max(1, len(c["x"]) // 6)
        c["x"][:m] = np.asarray(c["lon"], dtype=np.float64).ravel()[rng.integers(0, c["lon"].size, m)] + rng.normal(0, 0.3 * cell, m)
        c["y"][:m] = np.asarray(c["lat"], dtype=np.float64).ravel()[rng.integers(0, c["lat"].size, m)] + rng.normal(0, 0.3 * cell, m)
    tmax = float(c["times"][-1])
    if dt < 0:
        c["t"] = np.full(len(c["x"]), tmax)
    total = min(nstep * abs(dt), tmax)
    segs = [total] if rng.random() < 0.6 else [abs(dt) * max(1, nstep // 2), abs(dt) * max(1, nstep - nstep // 2)]
    segs = [s for s in segs if s > 0]
    if sum(segs) > tmax:
        segs = [tmax]
    return spec, c, kern, dt, segs


def compare_traj(d, pd, c, tol):
    if len(d["x"]) != len(pd["x"]):
        return [f"survivors {len(d['x'])} vs {len(pd['x'])}"]
    msg = [k for k in ("particle_id", "state", "t", "ei") if not np.array_equal(d[k], pd[k])]
    # C-grid velocities are discontinuous across cell faces: on a rough random field a last-bit difference (DESIGN.md waivers 8, 9) of
    # the one or two particles that sit on a face can grow over a long run before it decays again (traced: <= 1.4e-5 of a cell) --
    # beyond the ulp tolerance a C-grid trajectory is held to 1e-4 of a cell
    cell = None
    if c.get("interp", "cgrid_velocity") == "cgrid_velocity" and np.ndim(c["lon"]) == 2:
        cell = float(np.abs(np.diff(np.asarray(c["lon"], dtype=np.float64), axis=1)).mean())
    for key in "xyz":
        floor = 0.01 * float(np.abs(np.asarray(c[key])).max()) or None
        u = ulp_diff_f32(d[key], pd[key], floor=floor)
        if u.size and u.max() > tol:
            if cell is not None and key in "xy" and np.abs(d[key].astype(np.float64) - pd[key].astype(np.float64)).max() <= 1e-4 * cell:
                continue
            msg.append(f"{key}: {u.max():.1f} ulp")
    return msg


def fuzz_curv(rng):
    spec, c, kern, dt, segs = curv_case(rng)
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    ofs = oracle_fieldset(c)
    # Pre-populated cell guesses (ParticleSet.populate_indices, reference particleset.py:252-262): the first evaluation is then
    # hinted.  A FRESH set's first evaluation goes through the spatial hash for the whole batch and the reference computes that one
    # evaluation with float32-TYPED barycentric coordinates (DESIGN.md waiver 4: 1e-7 relative in that evaluation) -- in a rough
    # random field that seed grows along the trajectory, so fresh sets are compared with the tolerance of the waiver, below.
    fresh = bool(rng.random() < 0.3)
    if not fresh:
        ps.populate_indices()
        (zi, _), (yi, _), (xi, _) = po.grid_search(ofs.grid, pd["z"], pd["y"], pd["x"])
        pd["ei"][:, 0] = po.ravel_index(ofs.grid, zi, yi, xi)
        if not np.array_equal(ps._data["ei"], pd["ei"]):
            return f"curv {spec['interp']} populate_indices", spec, ["ei after populate_indices"]
    for s in segs:
        ps.execute([getattr(pb, kern), pb.DeleteParticle], dt=dt, runtime=s)
        po.pset_execute(pd, ofs, [getattr(po, kern), po.DeleteOnError], dt, runtime=s)
    what = f"curv {spec['interp']} {kern} {spec['mesh']} {spec['cdtype']} dt={dt} segs={segs} fresh={fresh}"
    if rng.random() < 0.3 and c["U"].shape[0] >= 3:
        # the same run with only 2 of the time levels resident (time-slab streaming behind the curvilinear search): bit-identical to
        # the resident engine run.  Steps that straddle a level sample three levels: such windows are refused, not wrong.
        wfs = pb.FieldSet.from_arrays(lon=c["lon"], lat=c["lat"], depth=c["depth"], time=c["times"], U=c["U"], V=c["V"], W=c["W"],
                                      mesh=c["mesh"], interp_method=c.get("interp", "cgrid_velocity"),
                                      padding=c.get("padding", ("low", "low", "high")), time_window=2)  # fmt: skip
        wps = pb.ParticleSet(wfs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
        if not fresh:
            wps.populate_indices()
        try:
            for s in segs:
                wps.execute([getattr(pb, kern), pb.DeleteParticle], dt=dt, runtime=s)
        except RuntimeError as e:
            if "cannot cover one step" not in str(e):
                raise
        else:
            what += " +window"
            diff = [k for k in ("particle_id", "state", "t", "ei", "x", "y", "z") if not np.array_equal(wps._data[k], ps._data[k])]
            if diff:
                return what, spec, [f"time-windowed run differs from the resident one: {diff}"]
        wfs.release()
    if fresh:  # ids / states / times exact; positions within 1e-4 of a cell (the float32-typed first evaluation, grown over the run)
        d = ps._data
        if len(d["x"]) != len(pd["x"]):
            return what, spec, [f"survivors {len(d['x'])} vs {len(pd['x'])}"]
        msg = [k for k in ("particle_id", "state", "t") if not np.array_equal(d[k], pd[k])]
        cell = float(np.abs(np.diff(np.asarray(c["lon"], dtype=np.float64), axis=1)).mean())
        for key in "xy":
            if len(d[key]) and np.abs(d[key].astype(np.float64) - pd[key].astype(np.float64)).max() > 1e-4 * cell:
                msg.append(f"{key}: {np.abs(d[key].astype(np.float64) - pd[key].astype(np.float64)).max() / cell:.2e} cells")
        return what, spec, msg
    return what, spec, compare_traj(ps._data, pd, c, CURV_ULP)


def fuzz_curv_scalar(rng):
    spec, c, kern, dt, segs = curv_case(rng, interp="cgrid_velocity")
    how = str(rng.choice(["linear", "nearest", "cgrid_tracer"]))
    T = c["U"].shape[0] if rng.random() < 0.6 else 1
    P = (1.0 + rng.uniform(0, 1, (T,) + c["U"].shape[1:])).astype(rng.choice([np.float32, np.float64]))
    fs = make_fieldset(c)
    fs.add_field("P", P, interp_method=how)
    n = len(c["x"])
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    tq = rng.uniform(0, float(c["times"][-1]), n) if T > 1 else np.zeros(n)
    v1 = fs.P.eval(tq, ps._data["z"], ps._data["y"], ps._data["x"], ps)
    cell = float(np.abs(np.diff(np.asarray(c["lon"], dtype=np.float64), axis=1)).mean())
    x2 = np.asarray(ps._data["x"], dtype=np.float64) + 0.3 * cell
    v2 = fs.P.eval(tq, ps._data["z"], np.asarray(ps._data["y"], dtype=np.float64), x2, ps)
    ofs = oracle_fieldset(c)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    view = po.View(pd, np.ones(n, dtype=bool))
    o1 = po.eval_scalar(ofs, P, how, tq, pd["z"], pd["y"], pd["x"], view)
    o2 = po.eval_scalar(ofs, P, how, tq, pd["z"], np.asarray(pd["y"], dtype=np.float64), x2, view)
    msg = []
    if not np.array_equal(ps._data["ei"], pd["ei"]):
        msg.append("ei")
    if not np.array_equal(ps._data["state"], pd["state"]):
        msg.append("state")
    tol = (64 * np.finfo(np.float32).eps * float(np.abs(P).max())) if how == "linear" else 0.0
    for a, b, nm in ((v1, o1, "value"), (v2, o2, "value2")):
        if np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max(initial=0.0) > tol:
            msg.append(nm)
    return f"curv scalar {how} {spec['mesh']} {spec['cdtype']}", spec, msg


def fuzz_advdiff_cgrid(rng):
    mesh = str(rng.choice(["flat", "spherical"]))
    spec = dict(seed=int(rng.integers(1, 10**6)), kind="smooth", interp="cgrid_velocity", cdtype=str(rng.choice(["f4", "f8"])),
                ddtype=str(rng.choice(["f4", "f8"])), mesh=mesh, nx=int(rng.integers(6, 22)), ny=int(rng.integers(6, 18)),
                nz=int(rng.integers(2, 5)), nt=int(rng.choice([2, 4])), tstep=float(rng.choice([500.0, 3600.0])), n=int(rng.integers(1, 100)),
                kernels=["AdvectionRK4"], dt=100.0, segments=[dict(runtime=100.0)], delete=True, margin=float(rng.choice([0.05, 0.2])),
                umax=float(rng.choice([0.5, 3.0])))  # fmt: skip
    c = cases.build(spec)
    c["W"] = None
    c["z"] = np.abs(np.asarray(c["z"]))
    kern = str(rng.choice(["AdvectionDiffusionM1", "AdvectionDiffusionEM"]))
    ktime = bool(rng.random() < 0.5)
    kd = rng.choice([np.float32, np.float64])
    shape = ((c["U"].shape[0] if ktime else 1),) + c["U"].shape[1:]
    kscale = 10.0 if mesh == "flat" else 100.0
    kz = (kscale * (1.0 + rng.uniform(0, 1, shape))).astype(kd)
    km = (kscale * (1.0 + rng.uniform(0, 1, shape))).astype(kd)
    dres = float(np.float64(c["lon"][1]) - np.float64(c["lon"][0])) * float(rng.choice([0.5, 1.0]))
    dt = float(rng.choice([20.0, 100.0])) * (1 if rng.random() < 0.8 else -1)
    runtime = abs(dt) * int(rng.integers(2, 8))
    if dt < 0:
        c["t"] = np.full(len(c["x"]), float(c["times"][-1]))
    runtime = min(runtime, float(c["times"][-1]))
    seed = int(rng.integers(1, 10**6))
    fs = make_fieldset(c)
    fs.add_field("Kh_zonal", kz)
    fs.add_field("Kh_meridional", km)
    fs.add_context("dres", dres)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"], seed=seed)
    ps.execute([getattr(pb, kern), pb.DeleteParticle], dt=dt, runtime=runtime)
    ofs = oracle_fieldset(c)
    ofs.scalars = {"Kh_zonal": (kz, "linear"), "Kh_meridional": (km, "linear")}
    ofs.context["dres"] = dres
    st = {"it": 0}

    def normal(view):
        zx, zy = device_normals(seed, 1, st["it"], view.particle_id)
        st["it"] += 1
        return zx, zy

    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    po.pset_execute(pd, ofs, [getattr(po, kern)(normal), po.DeleteOnError], dt, runtime=runtime)
    return f"advdiff cgrid {kern} {mesh} dt={dt}", spec, compare_traj(ps._data, pd, c, 0 if mesh == "flat" else 4)


def fuzz_migration(rng):
    world = int(rng.integers(2, 6))
    n = int(rng.integers(50, 1500))
    f = dict(zip(("lon", "lat", "depth", "times"), (np.linspace(-40.0, 40.0, int(rng.integers(30, 70))), np.linspace(-30.0, 30.0, 20),
                                                    800.0 * np.linspace(0, 1, 5) ** 1.5, np.arange(3) * 3600.0)))  # fmt: skip
    shape = (3, 5, 20, len(f["lon"]))
    U = (40.0 * rng.uniform(-1, 1, shape)).astype(np.float32)
    V = (20.0 * rng.uniform(-1, 1, shape)).astype(np.float32)
    W = (1e-3 * rng.uniform(-1, 1, shape)).astype(np.float32)
    x, y, z = rng.uniform(-38, 38, n), rng.uniform(-28, 28, n), rng.uniform(5, 700, n)
    dt, runtime = 600.0, 600.0 * int(rng.integers(3, 12))
    kernels = [pb.AdvectionRK4_3D, pb.DeleteParticle]
    full = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=x, y=y, z=z, t=np.zeros(n), particle_id=np.arange(n)))
    slabs = [D.DecomposedFieldSet(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["times"], U=U, V=V, W=W, mesh="spherical",
                                  rank=r, world=world, halo_cells=3, device=0) for r in range(world)]  # fmt: skip
    cap = int(rng.choice([4, 50, n + 8]))
    bases = [s.engine.migrate_p2p_init(cap)[1] for s in slabs]
    for s in slabs:
        s.engine.migrate_p2p_connect(local_bases=bases)
    plans = [D.decomposed_plan(s, kernels) for s in slabs]
    for r, s in enumerate(slabs):
        D.upload_decomposed(s, D.shard_particles(full, r, world), dt)
    rounds, halo = 0, False
    while True:
        reps = [s.engine.advect(D._advect_args(s.engine, plans[r], dt, runtime, rounds == 0, 0, 1, rounds)) for r, s in enumerate(slabs)]
        for s in slabs:
            s.engine.migrate_p2p_finish()
        rounds += 1
        halo = halo or any(rp["max_state"] == 99 for rp in reps)
        if halo or sum(rp["n_migrate"] for rp in reps) == 0 or rounds > 5000:
            break
    outs = [D.download_decomposed(s, dt) for s in slabs]
    for s in slabs:
        s.fs.release()
    if halo:
        return f"migration world={world} (halo violation: skipped)", None, []
    merged = {k: np.concatenate([o[k] for o in outs], axis=0) for k in outs[0]}
    order = np.argsort(merged["particle_id"], kind="stable")
    merged = {k: v[order] for k, v in merged.items()}
    fs = pb.FieldSet.from_arrays(lon=f["lon"], lat=f["lat"], depth=f["depth"], time=f["times"], U=U, V=V, W=W, mesh="spherical")
    ps = pb.ParticleSet(fs, x=x, y=y, z=z, t=np.zeros(n))
    ps.execute(kernels, dt=dt, runtime=runtime)
    ref = ps._data
    fs.release()
    msg = [k for k in ("particle_id", "state", "t", "ei", "x", "y", "z") if not (merged[k].shape == ref[k].shape and np.array_equal(merged[k], ref[k]))]
    if rounds > 5000:
        msg.append("no termination")
    return f"migration world={world} n={n} inbox={cap} rounds={rounds}", dict(world=world, n=n, cap=cap), msg


def fuzz_migration_windowed(rng):
    """Time-slab streaming under mode D through the product's own loop (execute_decomposed -> run_decomposed_p2p) on thread ranks:
    random slab count, window, level count, time direction, inbox capacity and staggered releases, against the undecomposed run
    with every level resident."""
    import thread_ranks

    thread_ranks.serialise_engine_calls()
    world = int(rng.integers(2, 5))
    n = int(rng.integers(50, 900))
    nt = int(rng.integers(3, 8))
    window = int(rng.integers(2, min(nt - 1, 4) + 1))  # (window + 1 prefetch slot <= levels of the field)
    tstep = 3600.0
    times = np.arange(nt) * tstep
    nx = int(rng.integers(30, 60))
    lon, lat, depth = np.linspace(-40.0, 40.0, nx), np.linspace(-30.0, 30.0, 20), 800.0 * np.linspace(0, 1, 5) ** 1.5
    shape = (nt, 5, 20, nx)
    U = (40.0 * rng.uniform(-1, 1, shape)).astype(np.float32)
    V = (20.0 * rng.uniform(-1, 1, shape)).astype(np.float32)
    W = (1e-3 * rng.uniform(-1, 1, shape)).astype(np.float32)
    x, y, z = rng.uniform(-38, 38, n), rng.uniform(-28, 28, n), rng.uniform(5, 700, n)
    sign = 1 if rng.random() < 0.65 else -1
    # level-aligned steps when the window is 2 (a step that straddles a level samples 3 levels)
    dt = sign * (float(rng.choice([600.0, 900.0, 1200.0])) if window == 2 else float(rng.choice([600.0, 700.0, 1100.0])))
    t_start = 0.0 if sign > 0 else float(times[-1])
    span = float(times[-1])
    t = t_start + sign * rng.choice([0.0, 0.0, abs(dt) * 3, tstep, 2 * tstep + abs(dt)], n)
    t = np.clip(t, 0.0, span)
    endtime = t_start + sign * float(rng.uniform(0.3, 1.0)) * span
    endtime = t_start + sign * abs(dt) * max(1, round(abs(endtime - t_start) / abs(dt))) if window == 2 else endtime
    endtime = float(np.clip(endtime, 0.0, span))
    kernels = [pb.AdvectionRK4_3D, pb.DeleteParticle]
    full = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=x, y=y, z=z, t=t, particle_id=np.arange(n)))
    slabs = [D.DecomposedFieldSet(lon=lon, lat=lat, depth=depth, time=times, U=U, V=V, W=W, mesh="spherical", rank=r, world=world,
                                  halo_cells=3, device=0, time_window=window) for r in range(world)]  # fmt: skip
    cap = int(rng.choice([4, 50, n + 8]))
    bases = [s.engine.migrate_p2p_init(cap)[1] for s in slabs]
    for s in slabs:
        s.engine.migrate_p2p_connect(local_bases=bases)
        s.p2p = True
    what = f"windowed migration world={world} n={n} nt={nt} window={window} dt={dt} inbox={cap}"
    try:
        res = thread_ranks.run_ranks(world, lambda r, dist: D.execute_decomposed(slabs[r], D.shard_particles(full, r, world), kernels, dt,
                                                                                 endtime, dist))  # fmt: skip
    except RuntimeError as e:
        if "halo violation" in str(e):
            return what + " (halo violation: skipped)", None, []
        raise
    finally:
        for s in slabs:
            s.fs.release()
    outs = [o for o, _ in res]
    merged = {k: np.concatenate([o[k] for o in outs], axis=0) for k in outs[0]}
    order = np.argsort(merged["particle_id"], kind="stable")
    merged = {k: v[order] for k, v in merged.items()}
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, depth=depth, time=times, U=U, V=V, W=W, mesh="spherical")
    ps = pb.ParticleSet(fs, x=x, y=y, z=z, t=t)
    ps.execute(kernels, dt=dt, endtime=endtime)
    ref = ps._data
    fs.release()
    msg = [k for k in ("particle_id", "state", "t", "ei", "x", "y", "z") if not (merged[k].shape == ref[k].shape and np.array_equal(merged[k], ref[k]))]
    return what + f" rounds={res[0][1]['rounds']}", dict(world=world, n=n, nt=nt, window=window, dt=dt, cap=cap), msg


def fuzz_multigrid(rng):
    spec = dict(seed=int(rng.integers(1, 10**6)), kind="smooth", cdtype="f8", ddtype="f4", mesh="spherical", nx=int(rng.integers(8, 20)),
                ny=int(rng.integers(8, 18)), nz=4, nt=3, tstep=3600.0, n=int(rng.integers(1, 80)), kernels=["AdvectionRK4_3D"], dt=600.0,
                segments=[dict(runtime=600.0)], delete=True, margin=0.1, umax=1.0)  # fmt: skip
    c = cases.build(spec)
    fs = make_fieldset(c)
    lon, lat = np.asarray(c["lon"], dtype=np.float64), np.asarray(c["lat"], dtype=np.float64)
    g2 = XGrid(np.linspace(lon[0] - 1, lon[-1] + 1, int(rng.integers(4, 12))), np.linspace(lat[0] - 1, lat[-1] + 1, int(rng.integers(4, 10))), None,
               mesh="spherical")  # fmt: skip
    T = 3 if rng.random() < 0.6 else 1
    how = str(rng.choice(["linear", "nearest"]))
    P = rng.uniform(-5, 5, (T, 1, len(g2.lat), len(g2.lon))).astype(rng.choice([np.float32, np.float64]))
    fs.add_field("P2", P, grid=g2, interp_method=how)
    U2 = rng.uniform(-8, 8, (3, 1, len(g2.lat), len(g2.lon))).astype(np.float32)
    V2 = rng.uniform(-8, 8, (3, 1, len(g2.lat), len(g2.lon))).astype(np.float32)
    g3 = XGrid(g2.lon + 0.3, g2.lat - 0.2, None, mesh="spherical")
    fs.add_vector_field("wind", U2, V2, grid=g3)
    n = len(c["x"])
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    tq = rng.uniform(0, 7200.0, n)
    v = fs.P2.eval(tq, ps._data["z"], ps._data["y"], ps._data["x"], ps)
    ei_after_p = ps._data["ei"][:, -1].copy()
    u, w = fs.wind.eval(tq, ps._data["z"], ps._data["y"], ps._data["x"], ps)
    times = np.asarray(c["times"], dtype=np.float64)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"], ngrids=3)
    view = po.View(pd, np.ones(n, dtype=bool))
    ofs2 = po.OFieldSet(po.OGrid(g2.lon, g2.lat, None, mesh="spherical"), P, P, None, time=times if T > 1 else None, interp="linear")
    ov = po.eval_scalar(ofs2, P, how, tq, pd["z"], pd["y"], pd["x"], view)
    oei = pd["ei"][:, -1].copy()
    ofs3 = po.OFieldSet(po.OGrid(g3.lon, g3.lat, None, mesh="spherical"), U2, V2, None, time=times, interp="linear")
    ou, ow = po.eval_uvw(ofs3, tq, pd["z"], pd["y"], pd["x"], view, False)
    msg = []
    if not np.array_equal(ei_after_p, oei) or not np.array_equal(ps._data["ei"][:, -1], pd["ei"][:, -1]):
        msg.append("ei")
    if not np.array_equal(ps._data["state"], pd["state"]):
        msg.append("state")
    if not np.array_equal(np.asarray(v, dtype=np.float64), np.asarray(ov, dtype=np.float64)):
        msg.append("scalar value")
    sc = max(float(np.abs(ou).max(initial=0.0)), 1e-30)
    if np.abs(u - ou).max(initial=0.0) > 4 * np.finfo(np.float32).eps * sc or np.abs(w - ow).max(initial=0.0) > 4 * np.finfo(np.float32).eps * sc:
        msg.append("vector value")
    fs.release()
    return f"multigrid {how} T={T}", spec, msg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    kinds = (fuzz_curv, fuzz_curv, fuzz_curv_scalar, fuzz_advdiff_cgrid, fuzz_migration, fuzz_multigrid, fuzz_migration_windowed)
    bad = 0
    for k in range(n):
        f = kinds[k % len(kinds)]
        try:
            what, spec, msg = f(rng)
        except Exception as e:  # noqa: BLE001
            import traceback
            what, spec, msg = f.__name__, None, [f"EXC {type(e).__name__}: {e}", traceback.format_exc(limit=4)]
        if msg:
            bad += 1
            print(f"[{k}] {what}: {'; '.join(map(str, msg))}\n    spec={spec}")
        elif os.environ.get("FUZZ_VERBOSE"):
            print(f"[{k}] ok: {what}")
    print(f"{n} cases, {bad} with differences")


if __name__ == "__main__":
    main()
