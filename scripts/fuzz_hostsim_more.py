"""More differential fuzzing of the host-compiled kernel sources (oracle/hostsim) against the oracle: scalar Field.eval
(XLinear / XNearest / CGrid_Tracer / XLinearInvdistLandTracer), AdvectionRK45, AdvectionDiffusionM1 / EM and fused
DiffusionUniformKh (same Philox normals), time-slab streaming vs the resident field.
Run:  PB_LIB=oracle/_build/hostsim/libparcels_b200_hostsim.so PB_HOSTSIM_TEST=1 python scripts/fuzz_hostsim_more.py [n] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import warnings
import numpy as np
import cases
import parcels_b200 as pb
from engine_run import make_fieldset, ulp_diff_f32
from oracle import parcels_oracle as po
from oracle_run import oracle_fieldset
from philox_ref import device_normals, wiener_normals

warnings.simplefilter("ignore")
ALL_INTERPS = ("linear", "linear", "freeslip", "partialslip", "nearest", "cgrid_velocity")


def base_case(rng, two_d, interps=("linear",)):
    mesh = str(rng.choice(["flat", "spherical"]))
    nt = int(rng.choice([1, 2, 4]))
    spec = dict(seed=int(rng.integers(1, 10**6)), kind="smooth", cdtype=str(rng.choice(["f4", "f8"])), ddtype=str(rng.choice(["f4", "f8"])),
                mesh=mesh, nx=int(rng.integers(5, 25)), ny=int(rng.integers(5, 20)), nz=int(rng.integers(2, 6)), nt=nt,
                tstep=float(rng.choice([500.0, 3600.0])), n=int(rng.integers(1, 120)), kernels=["AdvectionRK4" if two_d else "AdvectionRK4_3D"],
                dt=100.0, segments=[dict(runtime=100.0)], delete=True, margin=float(rng.choice([-0.03, 0.05, 0.2])),
                umax=float(rng.choice([0.5, 3.0])))  # fmt: skip
    how = str(rng.choice(list(interps)))
    if how == "nearest" and two_d:
        how = "linear"
    if how != "linear":
        spec["interp"] = how
    if how in ("freeslip", "partialslip"):
        spec["land"] = True
    c = cases.build(spec)
    if two_d:
        c["W"] = None
        c["z"] = np.abs(np.asarray(c["z"]))
    return spec, c


def fuzz_scalar(rng):
    spec, c = base_case(rng, two_d=False)
    T = c["U"].shape[0] if rng.random() < 0.6 else 1
    P = (1.0 + rng.uniform(0, 1, (T,) + c["U"].shape[1:])).astype(rng.choice([np.float32, np.float64]))
    P[rng.uniform(size=P.shape) < 0.15] = 0  # land
    how = str(rng.choice(["linear", "nearest", "cgrid_tracer", "linear_invdist_land"]))
    n = len(c["x"])
    if n == 1 and how == "linear_invdist_land":
        how = "linear"  # a batch of ONE sample: NumPy's sum over the corner axes becomes a contiguous pairwise sum (DESIGN.md waiver 1)
    tmax = 0.0 if c["times"] is None else float(c["times"][-1])
    tq = rng.uniform(0, tmax, n) if tmax else np.zeros(n)
    # (the land tracer SUMS over the gathered levels: a batch mixing on-level and off-level samples is waiver 1 of DESIGN.md)
    if rng.random() < 0.3 and n > 3 and how != "linear_invdist_land":  # some samples exactly on nodes / levels
        c["x"][:3] = np.asarray(c["lon"], dtype=np.float64)[rng.integers(0, len(c["lon"]), 3)]
        c["y"][:3] = np.asarray(c["lat"], dtype=np.float64)[rng.integers(0, len(c["lat"]), 3)]
        if c["times"] is not None:
            tq[:3] = c["times"][rng.integers(0, len(c["times"]), 3)]
    f32 = rng.random() < 0.5
    dt = np.float32 if f32 else np.float64
    x, y, z = (np.asarray(c[k]).astype(dt) for k in "xyz")
    fs = make_fieldset(c)
    fs.add_field("P", P, interp_method=how)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    val = fs.P.eval(tq, z, y, x, ps)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    oval = po.eval_scalar(oracle_fieldset(c), P, how, tq, z, y, x, po.View(pd, np.ones(n, dtype=bool)))
    msg = []
    # the reference decides lenT / lenZ per batch; per particle it differs only when the batch mixes tau == 0 / zeta == 0
    # particles with others (DESIGN.md waiver 1): those samples are compared on value only when both agree on the levels
    if val.dtype != oval.dtype and not (T > 1):
        msg.append(f"dtype {val.dtype} vs {oval.dtype}")
    if not np.array_equal(ps._data["ei"], pd["ei"]):
        msg.append("ei")
    if not np.array_equal(ps._data["state"], pd["state"]):
        msg.append("state")
    same = (val.astype(np.float64) == oval.astype(np.float64)) | (np.isnan(val) & np.isnan(oval))
    if not same.all():
        bad = np.flatnonzero(~same)
        msg.append(f"value at {bad[:4]}: {val[bad[:4]]} vs {oval[bad[:4]]}")
    return f"scalar {how} T={T} f32pos={f32} P={P.dtype}", spec, msg


def fuzz_rk45(rng):
    spec, c = base_case(rng, two_d=True, interps=("linear", "cgrid_velocity", "freeslip", "partialslip"))
    tmax = None if c["times"] is None else float(c["times"][-1])
    dt = float(rng.choice([50.0, 200.0])) * (1 if rng.random() < 0.8 else -1)
    nsteps = int(rng.integers(2, 10))
    runtime = abs(dt) * nsteps
    if tmax is not None:
        runtime = min(runtime, 0.45 * tmax)  # RK45 waiver: stay inside the time axis
        t0 = 0.0 if dt > 0 else tmax
        c["t"] = np.full(len(c["x"]), t0 + (0.05 * tmax if dt > 0 else -0.05 * tmax))
        if dt > 0 and rng.random() < 0.4:  # repeated release starting on the first time level: the reference's batch-level lenT
            c["t"] = np.where(rng.random(len(c["x"])) < 0.5, 0.0, 0.05 * tmax)
    tol = float(rng.choice([1e-4, 1e-2, 1.0]))
    min_dt, max_dt = float(rng.choice([0.5, 5.0])), abs(dt) * float(rng.choice([2, 4]))
    fs = make_fieldset(c)
    for k_, v_ in (("RK45_tol", tol), ("RK45_min_dt", min_dt), ("RK45_max_dt", max_dt)):
        fs.add_context(k_, v_)
    pclass = pb.Particle.add_variable(pb.Variable("next_dt", dtype=np.float32, initial=0))
    ps = pb.ParticleSet(fs, pclass=pclass, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    mixed = rng.random() < 0.4  # [AdvectionRK45, user kernel]: host loop control, RK45 attempts on the device

    def Drift(particles, fieldset):
        particles.dy += 0.25 * particles.dt

    ps.execute([pb.AdvectionRK45, Drift] if mixed else pb.AdvectionRK45, dt=dt, runtime=runtime)
    ofs = oracle_fieldset(c)
    ofs.context.update(RK45_tol=tol / ofs.grid.deg2m if ofs.grid.spherical else tol, RK45_min_dt=min_dt, RK45_max_dt=max_dt)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    pd["next_dt"] = np.zeros(len(pd["x"]), dtype=np.float32)

    def ODrift(p, fs_):
        p.dy = p.dy + 0.25 * p.dt

    po.pset_execute(pd, ofs, [po.AdvectionRK45, ODrift] if mixed else [po.AdvectionRK45], dt, runtime=runtime)
    msg = []
    d = ps._data
    for key in ("particle_id", "state", "t", "dt", "next_dt", "ei"):
        if not np.array_equal(d[key], pd[key]):
            msg.append(key)
    for key in "xy":
        floor = 0.01 * float(np.abs(np.asarray(c[key])).max()) or None
        u = ulp_diff_f32(d[key], pd[key], floor=floor)
        if u.size and u.max() > (0 if c["mesh"] == "flat" else 4):
            msg.append(f"{key}: {u.max():.1f} ulp")
    return f"rk45 tol={tol} dt={dt} mixed={mixed}", spec, msg


def fuzz_advdiff(rng):
    spec, c = base_case(rng, two_d=True)
    kern = str(rng.choice(["AdvectionDiffusionM1", "AdvectionDiffusionEM"]))
    ktime = c["U"].shape[0] > 1 and rng.random() < 0.6
    kd = rng.choice([np.float32, np.float64])
    shape = ((c["U"].shape[0] if ktime else 1),) + c["U"].shape[1:]
    scale = 40.0 if c["mesh"] == "flat" else 4000.0
    kz = (scale * (1.0 + rng.uniform(0, 1, shape))).astype(kd)
    km = (scale * (1.0 + rng.uniform(0, 1, shape))).astype(kd)
    dres = float(np.float64(c["lon"][1]) - np.float64(c["lon"][0])) * float(rng.choice([0.5, 1.0]))
    tmax = None if c["times"] is None else float(c["times"][-1])
    dt = float(rng.choice([20.0, 100.0])) * (1 if rng.random() < 0.8 else -1)
    runtime = abs(dt) * int(rng.integers(1, 8))
    if tmax is not None:
        runtime = min(runtime, 0.9 * tmax)
        c["t"] = np.full(len(c["x"]), 0.0 if dt > 0 else tmax)
        if dt > 0 and rng.random() < 0.4:  # repeated release starting on the first time level: the reference's batch-level lenT
            c["t"] = np.where(rng.random(len(c["x"])) < 0.5, 0.0, abs(dt) * float(rng.choice([0.5, 1.0])))
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
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    st = {"it": 0}

    def normal(view):
        zx, zy = device_normals(seed, 1, st["it"], view.particle_id)
        st["it"] += 1
        return zx, zy

    po.pset_execute(pd, ofs, [getattr(po, kern)(normal), po.DeleteOnError], dt, runtime=runtime)
    msg = []
    d = ps._data
    if len(d["x"]) != len(pd["x"]):
        return f"advdiff {kern}", spec, [f"survivors {len(d['x'])} vs {len(pd['x'])}"]
    for key in ("particle_id", "state", "t", "ei"):
        if not np.array_equal(d[key], pd[key]):
            msg.append(key)
    for key in "xy":
        start = np.asarray(c[key], dtype=np.float64)[pd["particle_id"]]
        floor = max(float(np.abs(pd[key] - start).max()) if len(start) else 0.0, 1e-30)
        u = ulp_diff_f32(d[key], pd[key], floor=floor)
        if u.size and u.max() > (1 if c["mesh"] == "flat" else 4):  # libm log / sincos vs the NumPy restatement of Box-Muller
            msg.append(f"{key}: {u.max():.1f} ulp")
    return f"advdiff {kern} ktime={ktime} kd={np.dtype(kd).name}", spec, msg


def fuzz_diffusion(rng):
    """fused [advection, DiffusionUniformKh, DeleteParticle] over one or two execute() segments, same Philox normals"""
    two_d = rng.random() < 0.5
    spec, c = base_case(rng, two_d=two_d, interps=ALL_INTERPS)
    c["constants"] = {"Kh_zonal": float(rng.choice([10.0, 100.0])), "Kh_meridional": float(rng.choice([5.0, 50.0]))}
    kern = str(rng.choice(["AdvectionRK4", "AdvectionEE", "AdvectionRK2"])) if two_d else str(rng.choice(["AdvectionRK4_3D", "AdvectionRK2_3D"]))
    tmax = None if c["times"] is None else float(c["times"][-1])
    dt = float(rng.choice([20.0, 100.0, 600.0])) * (1 if rng.random() < 0.8 else -1)
    nseg = int(rng.choice([1, 2, 3]))
    seg_rt = abs(dt) * int(rng.integers(1, 5))
    if tmax is not None:
        seg_rt = min(seg_rt, 0.9 * tmax / nseg)
        c["t"] = np.full(len(c["x"]), 0.0 if dt > 0 else tmax)
        if dt > 0 and rng.random() < 0.4:  # repeated release starting on the first time level: the reference's batch-level lenT
            c["t"] = np.where(rng.random(len(c["x"])) < 0.5, 0.0, abs(dt) * float(rng.choice([0.5, 1.0])))
    seed = int(rng.integers(1, 10**6))
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"], seed=seed)
    for _ in range(nseg):
        ps.execute([getattr(pb, kern), pb.DiffusionUniformKh, pb.DeleteParticle], dt=dt, runtime=seg_rt)
    ofs = oracle_fieldset(c)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"], ngrids=ofs.ngrids)
    st = {"call": 0, "it": 0}

    def normal(view):
        zx, zy = device_normals(seed, st["call"], st["it"], view.particle_id)
        st["it"] += 1
        return zx, zy

    for _ in range(nseg):
        st["call"] += 1
        st["it"] = 0
        po.pset_execute(pd, ofs, [getattr(po, kern), po.DiffusionUniformKh(normal), po.DeleteOnError], dt, runtime=seg_rt)
    msg = []
    d = ps._data
    if len(d["x"]) != len(pd["x"]):
        return f"diffusion {kern} x{nseg}", spec, [f"survivors {len(d['x'])} vs {len(pd['x'])}"]
    for key in ("particle_id", "state", "t", "ei"):
        if not np.array_equal(d[key], pd[key]):
            msg.append(key)
    for key in "xyz":
        start = np.asarray(c[key], dtype=np.float64)[pd["particle_id"]]
        floor = max(float(np.abs(pd[key] - start).max()) if len(start) else 0.0, 0.01 * float(np.abs(np.asarray(c[key])).max()), 1e-30)
        u = ulp_diff_f32(d[key], pd[key], floor=floor)
        if u.size and u.max() > (1 if c["mesh"] == "flat" else 4):
            msg.append(f"{key}: {u.max():.1f} ulp")
    return f"diffusion {kern} x{nseg}", spec, msg


def fuzz_window(rng):
    """time-slab streaming == fully resident field, bit for bit (advection only)"""
    spec, c = base_case(rng, two_d=False)
    nt = int(rng.choice([4, 6]))
    spec.update(nt=nt, tstep=600.0)
    c = cases.build(spec)
    tmax = float(c["times"][-1])
    window = int(rng.choice([2, 3]))
    sign = 1 if rng.random() < 0.7 else -1
    dt = sign * float(rng.choice([100.0, 200.0, 300.0, 600.0]))  # divides the 600 s level spacing: aligned steps (window 2 needs that)
    c["t"] = np.full(len(c["x"]), 0.0 if sign > 0 else tmax)
    runtime = float(rng.choice([0.5, 0.8, 1.0])) * tmax
    out = []
    for w in (None, window):
        fs = pb.FieldSet.from_arrays(lon=c["lon"], lat=c["lat"], depth=c["depth"], time=c["times"], U=c["U"], V=c["V"], W=c["W"],
                                     mesh=c["mesh"], time_window=w)  # fmt: skip
        ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
        ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=dt, runtime=runtime)
        out.append(ps._data)
    msg = [k for k in ("particle_id", "state", "t", "ei", "x", "y", "z") if not np.array_equal(out[0][k], out[1][k])]
    return f"window {window} dt={dt}", spec, msg


def fuzz_output(rng):
    """rows handed to the output file at every output time (device selection + compaction across resident intervals, deletions
    in HBM) == the rows the oracle's outer loop selects"""
    three = rng.random() < 0.5
    spec, c = base_case(rng, two_d=not three, interps=ALL_INTERPS)
    kern = "AdvectionRK4_3D" if three else "AdvectionRK4"
    tmax = None if c["times"] is None else float(c["times"][-1])
    dt = float(rng.choice([50.0, 100.0, 300.0])) * (1 if rng.random() < 0.8 else -1)
    nsteps = int(rng.integers(2, 14))
    runtime = abs(dt) * nsteps
    outputdt = abs(dt) * float(rng.choice([1, 2, 3, 2.5]))
    n = len(c["x"])
    if tmax is not None:
        runtime = min(runtime, 0.9 * tmax)
        lo = 0.0 if dt > 0 else tmax
        c["t"] = lo + np.sign(dt) * np.round(rng.uniform(0, 0.3 * runtime, n) / abs(dt)) * abs(dt) * (rng.random() < 0.5)
    delete = rng.random() < 0.8
    rows_e, rows_o = [], []

    class Rec:
        pass

    rec = Rec()
    rec.outputdt = outputdt
    rec.write = lambda pset, t: rows_e.append((float(t), pset._output_columns(float(t), ["t", "z", "y", "x", "particle_id"])[0]))
    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    err = oerr = ""
    try:
        ps.execute([getattr(pb, kern)] + ([pb.DeleteParticle] if delete else []), dt=dt, runtime=runtime, output_file=rec)
    except RuntimeError as e:
        if type(e).__module__.startswith("parcels_b200._lib"):
            raise
        err = type(e).__name__
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])

    def on_output(pdata, t):
        r = po.to_write_particles(pdata, t)
        rows_o.append((float(t), {k: pdata[k][r].copy() for k in ("t", "z", "y", "x", "particle_id")}))

    try:
        po.pset_execute(pd, oracle_fieldset(c), [getattr(po, kern)] + ([po.DeleteOnError] if delete else []), dt, runtime=runtime,
                        outputdt=outputdt, on_output=on_output)  # fmt: skip
    except po.OracleParticleError as e:
        oerr = str(e.code)
    msg = []
    if bool(err) != bool(oerr):
        msg.append(f"error {err!r} vs {oerr!r}")
    if len(rows_e) != len(rows_o):
        msg.append(f"{len(rows_e)} writes vs {len(rows_o)}")
    for (te, ce), (to, co) in zip(rows_e, rows_o):
        if te != to or not np.array_equal(ce["particle_id"], co["particle_id"]) or not np.array_equal(ce["t"], co["t"]):
            msg.append(f"rows at t={te}/{to}")
            break
        for key in "xyz":
            floor = 0.01 * float(np.abs(np.asarray(c[key])).max()) or None
            u = ulp_diff_f32(ce[key], co[key], floor=floor)
            if u.size and u.max() > (0 if c["mesh"] == "flat" else 4):
                msg.append(f"{key} at t={te}: {u.max():.1f} ulp")
    return f"output every {outputdt} dt={dt} delete={delete}", spec, msg


def fuzz_stepwise(rng):
    """mixed list [built-in, user kernel, user error handler]: host loop control + device kernels == oracle"""
    three = rng.random() < 0.5
    spec, c = base_case(rng, two_d=not three, interps=ALL_INTERPS)
    kern = str(rng.choice(["AdvectionRK4_3D", "AdvectionRK2_3D"])) if three else str(rng.choice(["AdvectionRK4", "AdvectionEE"]))
    tmax = None if c["times"] is None else float(c["times"][-1])
    dt = float(rng.choice([50.0, 100.0, 300.0])) * (1 if rng.random() < 0.8 else -1)
    runtime = abs(dt) * int(rng.integers(1, 8))
    if tmax is not None:
        runtime = min(runtime, 0.9 * tmax)
        c["t"] = np.full(len(c["x"]), 0.0 if dt > 0 else tmax)
        if dt > 0 and rng.random() < 0.4:  # repeated release starting on the first time level: the reference's batch-level lenT
            c["t"] = np.where(rng.random(len(c["x"])) < 0.5, 0.0, abs(dt) * float(rng.choice([0.5, 1.0])))
    lo, hi = float(c["lon"][1]), float(c["lon"][-2])

    def Periodic(particles, fieldset):
        xn = particles.x + particles.dx
        particles.dx += np.where(xn > hi, lo - hi, 0.0) + np.where(xn < lo, hi - lo, 0.0)

    def DeleteErr(particles, fieldset):
        particles[particles.state >= 50].state = 30

    fs = make_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    ps.execute([getattr(pb, kern), Periodic, DeleteErr], dt=dt, runtime=runtime)
    pd = po.create_particle_data(c["x"], c["y"], c["z"], c["t"])
    po.pset_execute(pd, oracle_fieldset(c), [getattr(po, kern), Periodic, po.DeleteOnError], dt, runtime=runtime)
    d = ps._data
    if len(d["x"]) != len(pd["x"]):
        return f"stepwise {kern}", spec, [f"survivors {len(d['x'])} vs {len(pd['x'])}"]
    msg = [k for k in ("particle_id", "state", "t", "ei") if not np.array_equal(d[k], pd[k])]
    for key in "xyz":
        floor = 0.01 * float(np.abs(np.asarray(c[key])).max()) or None
        u = ulp_diff_f32(d[key], pd[key], floor=floor)
        if u.size and u.max() > (0 if c["mesh"] == "flat" else 4):
            msg.append(f"{key}: {u.max():.1f} ulp")
    return f"stepwise {kern} dt={dt}", spec, msg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for k in range(n):
        f = (fuzz_scalar, fuzz_rk45, fuzz_advdiff, fuzz_diffusion, fuzz_window, fuzz_output, fuzz_stepwise)[k % 7]
        try:
            what, spec, msg = f(rng)
        except Exception as e:  # noqa: BLE001
            import traceback
            what, spec, msg = f.__name__, None, [f"EXC {type(e).__name__}: {e}", traceback.format_exc(limit=3)]
        if msg:
            bad += 1
            print(f"[{k}] {what}: {'; '.join(map(str, msg))}\n    spec={spec}")
    print(f"{n} cases, {bad} with differences")


if __name__ == "__main__":
    main()
