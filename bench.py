#!/usr/bin/env python
"""bench.py -- particle-RK4-steps/s of the hot path on N B200s (contract: see the task statement).

One bench "step" = one pass of the hot path over one batch: ``Kernel.execute`` of
AdvectionRK4_3D over the whole particle set for one output interval (144 dt-steps of 600 s = one
day on the config-2 field), i.e. 144 x N_particles particle-RK4-steps.

  value      whole-job particle-RK4-steps/s, particles + field already resident in HBM
  e2e        same metric through the public API ``ParticleSet.execute`` with HOST particle arrays:
             host->device upload and device->host download of the particle SoA inside the timed region
  roofline   algorithmic bytes (832 B per particle-RK4-step, SURVEY.md 8d / DESIGN.md) x steps per
             launch / CUDA-event duration of the advection kernel, against the measured HBM peak
  cpu_baseline  the oracle port (NumPy restatement of the reference path), one core, bounded sample

``--impl reference`` times the reference arm: the reference's own algorithm (oracle port; the
reference is pure Python and cannot travel to the GPU box) on the host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BYTES_PER_STEP = 832  # 4 stages x 3 comps x 16 corners x 4 B + 64 B particle state
METRIC = "particle-RK4-steps/sec"


# ------------------------------------------------------------------------------------------------
# workload: config 2 of BASELINE.json: rectilinear 1/4 deg x 50 levels, T=3, f32 U,V,W
# ------------------------------------------------------------------------------------------------
def c2_field(nx=1440, ny=720, nz=50, nt=3, seed=1):
    rng = np.random.default_rng(seed)
    lon = np.linspace(-180.0, 180.0, nx)
    lat = np.linspace(-80.0, 80.0, ny)
    depth = 5500.0 * (np.linspace(0.0, 1.0, nz) ** 1.8)
    times = np.arange(nt) * 86400.0
    X = (2 * np.pi * np.linspace(0, 1, nx)).astype(np.float32)[None, None, None, :]
    Y = (2 * np.pi * np.linspace(0, 1, ny)).astype(np.float32)[None, None, :, None]
    Z = np.linspace(0, 1, nz).astype(np.float32)[None, :, None, None]
    T = np.arange(nt, dtype=np.float32)[:, None, None, None]
    shape = (nt, nz, ny, nx)

    def noise():
        return rng.random(shape, dtype=np.float32) * np.float32(0.1) - np.float32(0.05)

    U = (np.sin(3 * X + 0.3 * T) * np.cos(2 * Y) * (1 - 0.5 * Z) * np.float32(0.6)
         + np.cos(5 * Y + T) * np.float32(0.3) + noise()).astype(np.float32)  # fmt: skip
    V = (np.cos(2 * X + 0.2 * T) * np.sin(4 * Y) * (1 - 0.3 * Z) * np.float32(0.6)
         + np.sin(3 * X) * np.float32(0.25) + noise()).astype(np.float32)  # fmt: skip
    W = ((np.sin(2 * X) * np.sin(3 * Y) * np.sin(np.float32(np.pi) * Z) * np.cos(0.5 * T) * np.float32(0.9)
          + noise()) * np.float32(1e-3)).astype(np.float32)  # fmt: skip
    return dict(lon=lon, lat=lat, depth=depth, times=times, U=U, V=V, W=W, mesh="spherical")


def c2_particles(n, seed):
    rng = np.random.default_rng(seed)
    return dict(x=rng.uniform(-170, 170, n), y=rng.uniform(-70, 70, n), z=rng.uniform(5, 5000, n), t=np.zeros(n))


WORKLOADS = {
    # name: (field kwargs, particles per GPU, dt, steps per pass)
    "c2": (dict(nx=1440, ny=720, nz=50, nt=3), 1_000_000, 600.0, 144),
    "c2_small": (dict(nx=360, ny=180, nz=20, nt=3), 100_000, 600.0, 144),  # quick functional check
}


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index=0):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()  # fmt: skip
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}  # fmt: skip


def measured_peak():
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(m["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def dram_traffic_per_launch(workload):
    """dram__bytes_read+write per launch from the committed ncu capture (profiles/), or None."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return j.get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference path)
# ------------------------------------------------------------------------------------------------
def _oracle_pass(field, parts, dt, nsteps):
    from oracle import parcels_oracle as po

    g = po.OGrid(field["lon"], field["lat"], field["depth"], mesh=field["mesh"])
    fs = po.OFieldSet(g, field["U"], field["V"], field["W"], time=field["times"])
    pd = po.create_particle_data(parts["x"], parts["y"], parts["z"], parts["t"])
    t0 = time.perf_counter()
    steps = po.pset_execute(pd, fs, [po.AdvectionRK4_3D, po.DeleteOnError], dt, runtime=dt * nsteps)
    return steps, time.perf_counter() - t0


_G = {}


def _worker(args):
    seed, n, dt, nsteps = args
    steps, _ = _oracle_pass(_G["field"], c2_particles(n, seed), dt, nsteps)
    return steps


def run_reference_arm(a, field, dt, nsteps, rank, world):
    """Reference arm: the reference's algorithm (oracle port) on all host cores, rank 0 only."""
    if rank != 0:
        return
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    per = a.ref_particles_per_core
    _G["field"] = field
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def one_pass(k):
            t0 = time.perf_counter()
            steps = sum(pool.map(_worker, [(1000 * k + c, per, dt, nsteps) for c in range(cores)]))
            return steps, time.perf_counter() - t0

        for k in range(a.warmup):
            one_pass(k)
        tot_steps, tot_t = 0, 0.0
        for k in range(a.steps):
            s, t = one_pass(100 + k)
            tot_steps += s
            tot_t += t
    v = tot_steps / tot_t
    sample = f"{cores} procs x {per} particles x {nsteps} dt-steps per bench step (same field, dt, kernel)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "particle-steps/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / max(a.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": a.workload, "kernel": "AdvectionRK4_3D", "note": "oracle port of the reference's NumPy path "
                   "(reference is pure Python and absent on the GPU box)"},
        "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="particles of the cpu_baseline sample")
    ap.add_argument("--ref-particles-per-core", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    fkw, n_per_gpu, dt, nsteps = WORKLOADS[a.workload]
    if a.particles:
        n_per_gpu = a.particles

    if a.impl == "reference":
        if rank == 0:
            run_reference_arm(a, c2_field(**fkw), dt, nsteps, rank, world)
        return

    import torch

    import parcels_b200 as pb
    from parcels_b200 import build

    build.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    field = c2_field(**fkw)
    fs = pb.FieldSet.from_arrays(lon=field["lon"], lat=field["lat"], depth=field["depth"], time=field["times"],
                                 U=field["U"], V=field["V"], W=field["W"], mesh="spherical")  # fmt: skip
    parts = c2_particles(n_per_gpu, seed=1 + rank)  # weak scaling: every rank owns its own 1e6-particle shard
    ps = pb.ParticleSet(fs, x=parts["x"], y=parts["y"], z=parts["z"], t=parts["t"], device=local_rank)
    init = {k: v.copy() for k, v in ps._data.items()}
    eng = fs.engine(local_rank)
    runtime = dt * nsteps
    kernels = [pb.AdvectionRK4_3D, pb.DeleteParticle]

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident arm: particles + fields in HBM; per step: restore snapshot + ONE kernel ----
    ei_last = np.ascontiguousarray(init["ei"][:, -1])
    eng.upload_particles(init, ei_last)
    eng.snapshot()
    args = eng.make_args(5, dt, runtime, delete_on_error=True)

    def resident_step():
        eng.restore()
        eng.advect_async(args)

    for _ in range(a.warmup):
        resident_step()
    barrier()
    ksum, psteps = 0.0, 0
    with ClockSampler(local_rank) as clk:
        eng.timer_begin()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            resident_step()
            rep = eng.last_report()  # waits for this step's kernel (report is read back every step)
            ksum += rep["kernel_ms"]
            psteps += rep["particle_steps"]
        dev_ms = eng.timer_end_ms()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = reduce_max(dev_ms)
    total_steps = reduce_sum(psteps)
    value = total_steps / (dev_ms * 1e-3)
    kernel_ms = ksum / a.steps
    steps_per_launch = psteps / a.steps
    refills = rep["cache_refills"]

    # ---- end-to-end arm: public API with host arrays; H2D + D2H of the particle SoA every step ----
    fresh = [{k: v.copy() for k, v in init.items()} for _ in range(a.steps)]  # host input batches, made before timing

    for _ in range(min(a.warmup, 3)):
        ps._data = {k: v.copy() for k, v in init.items()}
        ps.execute(kernels, dt=dt, runtime=runtime)
    barrier()
    e2e_steps = 0
    t0 = time.perf_counter()
    for i in range(a.steps):
        ps._data = fresh[i]
        ps.execute(kernels, dt=dt, runtime=runtime)
        e2e_steps += ps.last_report["particle_steps"]
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_value = reduce_sum(e2e_steps) / e2e_s
    n = n_per_gpu
    h2d = n * (6 * 4 + 8 + 4 + 4 + 8)
    d2h = n * (6 * 4 + 8 + 4 + 4)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_STEP * steps_per_launch / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{a.workload}: BASELINE.json configs[1] -- AdvectionRK4_3D, rectilinear {fkw['nx']}x{fkw['ny']}x{fkw['nz']} "
                        f"T={fkw['nt']} f32 U,V,W spherical, {n_per_gpu} particles/GPU, dt=600 s x {nsteps} steps per pass",
            "particles_per_gpu": n_per_gpu, "dt_steps_per_pass": nsteps,
            "l2_policy": "inputs larger than L2 (1.87 GB field, 52 MB particle SoA; particles re-seeded from an HBM snapshot every pass)",
            "wall_ms_per_step": wall_ms / a.steps, "kernel_ms_per_launch": kernel_ms, "corner_cache_refills_per_launch": refills,
        },
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "parcels_b200.ParticleSet.execute([AdvectionRK4_3D, DeleteParticle], dt=600, runtime=86400)"},
        "gpu_launches": a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": dram_traffic_per_launch(a.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_particle_step": BYTES_PER_STEP, "particle_steps_per_launch": steps_per_launch},
        "clocks": clk.summary(),
    }  # fmt: skip
    if not a.no_cpu_baseline:
        sample = c2_particles(a.cpu_sample, seed=1)
        s, t = _oracle_pass(field, sample, dt, nsteps)
        line["cpu_baseline"] = {"value": s / t, "unit": "particle-steps/s", "cores": 1, "kind": "port",
                                "sample": f"first {a.cpu_sample} particles of the same workload, {nsteps} dt-steps, "
                                          f"NumPy oracle port of the reference path ({t:.1f} s)"}  # fmt: skip
    print(json.dumps(line))


if __name__ == "__main__":
    main()
