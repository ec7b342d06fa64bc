#!/usr/bin/env python
"""bench.py -- particle-RK4-steps/s of the hot path on N B200s (contract: see the task statement).

One bench "step" = one pass of the hot path over one batch: ``Kernel.execute`` of the workload's
kernel list over the whole particle set for one output interval (e.g. 144 dt-steps of 600 s = one
day on the config-2 field), i.e. dt_steps x N_particles particle-RK4-steps.

  value      whole-job particle-RK4-steps/s, particles + field already resident in HBM
  e2e        same metric through the public API ``ParticleSet.execute`` with HOST particle arrays:
             host->device upload and device->host download of the particle SoA inside the timed region
  roofline   algorithmic bytes per particle-RK4-step (SURVEY.md 8d / DESIGN.md) x steps per launch /
             CUDA-event duration of the advection kernel, against the measured HBM peak
  cpu_baseline  the oracle port (NumPy restatement of the reference path), one core, bounded sample

Workloads (BASELINE.json configs): ``c2`` (default, configs[1]), ``c3`` (curvilinear C-grid, configs[2]),
``c4`` (RK4_3D + DiffusionUniformKh, configs[3]); ``*_small`` are quick functional variants.

``--impl reference`` times the reference arm: the reference's own algorithm (oracle port; the
reference is pure Python and cannot travel to the GPU box) on all host cores.
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

METRIC = "particle-RK4-steps/sec"
_G = {}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def c2_field(nx=1440, ny=720, nz=50, nt=3, seed=1):
    """config 2: rectilinear 1/4 deg x 50 levels, T=3 daily snapshots, f32 U,V,W, spherical."""
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


def ns_field_device(device=0, nx=4320, ny=2160, nz=50, nt=3, seed=4):
    """north-star target field (BASELINE.json north_star: "1e7 particles on a 1/12 deg 3D rectilinear field", 1 GPU):
    the config-2 formulas at 4320 x 2160 x 50, T=3 -- 5.6 GB per component, 16.8 GB in all -- generated IN HBM with
    torch and handed to the engine without a copy (FieldSet accepts __cuda_array_interface__ arrays)."""
    import torch

    dev = torch.device(f"cuda:{device}")
    g = torch.Generator(device=dev).manual_seed(seed)
    lon = np.linspace(-180.0, 180.0, nx)
    lat = np.linspace(-80.0, 80.0, ny)
    depth = 5500.0 * (np.linspace(0.0, 1.0, nz) ** 1.8)
    times = np.arange(nt) * 86400.0
    X = (2 * np.pi * torch.linspace(0, 1, nx, device=dev))[None, None, :]
    Y = (2 * np.pi * torch.linspace(0, 1, ny, device=dev))[None, :, None]
    Z = torch.linspace(0, 1, nz, device=dev)[:, None, None]
    U, V, W = (torch.empty((nt, nz, ny, nx), dtype=torch.float32, device=dev) for _ in range(3))

    def noise():
        return torch.rand((nz, ny, nx), generator=g, device=dev, dtype=torch.float32) * 0.1 - 0.05

    for k in range(nt):
        T = float(k)
        U[k] = torch.sin(3 * X + 0.3 * T) * torch.cos(2 * Y) * (1 - 0.5 * Z) * 0.6 + torch.cos(5 * Y + T) * 0.3 + noise()
        V[k] = torch.cos(2 * X + 0.2 * T) * torch.sin(4 * Y) * (1 - 0.3 * Z) * 0.6 + torch.sin(3 * X) * 0.25 + noise()
        W[k] = (torch.sin(2 * X) * torch.sin(3 * Y) * torch.sin(np.pi * Z) * float(np.cos(0.5 * T)) * 0.9 + noise()) * 1e-3
    torch.cuda.synchronize(dev)
    return dict(lon=lon, lat=lat, depth=depth, times=times, U=U, V=V, W=W, mesh="spherical")


def c2_particles(field, n, seed):
    rng = np.random.default_rng(seed)
    return dict(x=rng.uniform(-170, 170, n), y=rng.uniform(-70, 70, n), z=rng.uniform(5, 5000, n), t=np.zeros(n))


def c3_field(nx=1442, ny=1021, nt=3, seed=2):
    """config 3: curvilinear C-grid of ORCA025 shape (ny, nx) = (1021, 1442): rotated-pole mesh with a
    tanh-stretched latitude, f32 node coordinates (NEMO style), NEMO staggering (offsets X=1, Y=1), 2-D."""
    rng = np.random.default_rng(seed)
    lam = np.deg2rad(np.linspace(-70.0, 70.0, nx))[None, :]          # rotated longitude
    s = np.linspace(-1.0, 1.0, ny)[:, None]
    phi = np.deg2rad(42.0 * np.tanh(1.2 * s) / np.tanh(1.2))         # rotated latitude, finer near the "equator"
    pole = np.deg2rad(25.0)                                          # tilt of the rotated pole
    xr, yr, zr = np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi) * np.ones_like(lam)
    xg = np.cos(pole) * xr - np.sin(pole) * zr
    zg = np.sin(pole) * xr + np.cos(pole) * zr
    lon = np.rad2deg(np.arctan2(yr, xg)).astype(np.float32)
    lat = np.rad2deg(np.arcsin(np.clip(zg, -1, 1))).astype(np.float32)
    times = np.arange(nt) * 86400.0
    I = np.linspace(0, 2 * np.pi, nx, dtype=np.float32)[None, None, None, :]
    J = np.linspace(0, 2 * np.pi, ny, dtype=np.float32)[None, None, :, None]
    T = np.arange(nt, dtype=np.float32)[:, None, None, None]
    shape = (nt, 1, ny, nx)

    def noise():
        return rng.random(shape, dtype=np.float32) * np.float32(0.1) - np.float32(0.05)

    U = (np.float32(0.5) * np.sin(3 * I + 0.4 * T) * np.cos(2 * J) + np.float32(0.2) * np.cos(4 * J) + noise()).astype(np.float32)
    V = (np.float32(0.5) * np.cos(2 * I) * np.sin(3 * J + 0.3 * T) + np.float32(0.2) * np.sin(5 * I) + noise()).astype(np.float32)
    return dict(lon=lon, lat=lat, depth=None, times=times, U=U, V=V, W=None, mesh="spherical", interp="cgrid_velocity",
                padding=("low", "low", "high"))  # fmt: skip


def c3_particles(field, n, seed):
    """uniform over the interior of the mesh: random cell + bilinear blend of its corners"""
    rng = np.random.default_rng(seed)
    lon, lat = field["lon"].astype(np.float64), field["lat"].astype(np.float64)
    ny, nx = lon.shape
    jj, ii = rng.uniform(8, ny - 9, n), rng.uniform(8, nx - 9, n)
    j0, i0 = jj.astype(np.int64), ii.astype(np.int64)
    fj, fi = jj - j0, ii - i0

    def bl(a):
        return ((1 - fj) * (1 - fi) * a[j0, i0] + (1 - fj) * fi * a[j0, i0 + 1] + fj * fi * a[j0 + 1, i0 + 1]
                + fj * (1 - fi) * a[j0 + 1, i0])  # fmt: skip

    return dict(x=bl(lon), y=bl(lat), z=np.zeros(n), t=np.zeros(n))


def _hash_noise(t, z, y, x, salt):
    """deterministic noise in [-0.05, 0.05) from GLOBAL indices (so neighbouring slabs agree on halo columns)"""
    h = np.sin(t * 12.9898 + z * 78.233 + y * 37.719 + x * 4.1414 + salt) * 43758.5453
    return ((h - np.floor(h)) * 0.1 - 0.05).astype(np.float32)


def c5_slab(rank, world, halo, nx=4320, ny=2160, nz=50, nt=2):
    """config 5: rectilinear 1/12 deg x 50 levels field, THIS RANK'S X-slab only (columns lo..hi of the global
    axis incl. halo); analytic modes + index-hashed noise, so every rank can build its slab independently."""
    from parcels_b200.distributed import slab_plan

    lon = np.linspace(-180.0, 180.0, nx)
    lat = np.linspace(-80.0, 80.0, ny)
    depth = 5500.0 * (np.linspace(0.0, 1.0, nz) ** 1.8)
    times = np.arange(nt) * 86400.0
    plan = slab_plan(lon, world, halo)[rank]
    lo, hi = plan["lo"], plan["hi"]
    xi = np.arange(lo, hi + 1, dtype=np.float64)[None, None, None, :]
    yi = np.arange(ny, dtype=np.float64)[None, None, :, None]
    zi = np.arange(nz, dtype=np.float64)[None, :, None, None]
    ti = np.arange(nt, dtype=np.float64)[:, None, None, None]
    X, Y, Z = 2 * np.pi * xi / (nx - 1), 2 * np.pi * yi / (ny - 1), zi / (nz - 1)
    U = (np.sin(3 * X + 0.3 * ti) * np.cos(2 * Y) * (1 - 0.5 * Z) * 0.6 + np.cos(5 * Y + ti) * 0.3).astype(np.float32)
    U = U + _hash_noise(ti, zi, yi, xi, 0.1)
    V = (np.cos(2 * X + 0.2 * ti) * np.sin(4 * Y) * (1 - 0.3 * Z) * 0.6 + np.sin(3 * X) * 0.25).astype(np.float32)
    V = V + _hash_noise(ti, zi, yi, xi, 1.7) + np.zeros_like(U)
    W = ((np.sin(2 * X) * np.sin(3 * Y) * np.sin(np.pi * Z) * np.cos(0.5 * ti) * 0.9).astype(np.float32)
         + _hash_noise(ti, zi, yi, xi, 2.9)) * np.float32(1e-3)  # fmt: skip
    return dict(lon=lon, lat=lat, depth=depth, times=times, U=np.ascontiguousarray(U), V=np.ascontiguousarray(V),
                W=np.ascontiguousarray(W.astype(np.float32)), lo=lo, hi=hi, plan=plan)


def run_decomposed_bench(a, rank, local_rank, world):
    """Mode D bench (config 5): field cut into X-slabs, particles migrate over NCCL.  One bench step = one
    Kernel.execute over the decomposed field (advect kernels + migration rounds + all-to-all-v)."""
    import torch
    import torch.distributed as dist

    import parcels_b200 as pb
    from parcels_b200 import build
    from parcels_b200 import distributed as D
    from parcels_b200.particle import create_particle_data

    build.build()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    small = a.workload.endswith("_small")
    dims = dict(nx=432, ny=216, nz=20, nt=2) if small else dict(nx=4320, ny=2160, nz=50, nt=2)
    n_per_gpu = a.particles or (200_000 if small else 12_500_000)
    dt, nsteps, halo = 600.0, 48, 3
    f = c5_slab(rank, world, halo, **dims)
    fs = pb.FieldSet.from_arrays(lon=f["lon"][f["lo"] : f["hi"] + 1].copy(), lat=f["lat"], depth=f["depth"], time=f["times"],
                                 U=f["U"], V=f["V"], W=f["W"], mesh="spherical", xdim=f["lon"].size - 1)  # fmt: skip
    dfs = D.DecomposedFieldSet.from_slab(fs, f["plan"], rank=rank, world=world, device=local_rank)
    rng = np.random.default_rng(100 + rank)
    b = f["plan"]["bounds"]
    # every rank seeds its particles inside its own slab (as a domain-decomposed application would); whatever
    # leaves the slab during the pass migrates over NCCL
    n = n_per_gpu
    x = rng.uniform(max(b[rank], -175.0), min(b[rank + 1], 175.0), n)
    y, z = rng.uniform(-70, 70, n), rng.uniform(5, 5000, n)
    pid = np.arange(n, dtype=np.int64) + rank * n
    runtime = dt * nsteps
    kernels = [pb.AdvectionRK4_3D, pb.DeleteParticle]
    dev = f"cuda:{local_rank}"

    def one_pass():
        pdata = create_particle_data(nparticles=n, ngrids=1, initial=dict(x=x, y=y, z=z, t=np.zeros(n), particle_id=pid))  # input batch
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        out, stats = D.execute_decomposed(dfs, pdata, kernels, dt, runtime, dist)
        torch.cuda.synchronize()
        dist.barrier()
        return time.perf_counter() - t0, stats

    for _ in range(a.warmup):
        one_pass()
    tot_t, tot_steps, tot_mig, rounds, kms = 0.0, 0, 0, 0, 0.0
    with ClockSampler(local_rank) as clk:
        for _ in range(a.steps):
            t, st = one_pass()
            tot_t += D.allreduce_max(t, dist, dev)
            tot_steps += D.allreduce_sum(st["particle_steps"], dist, dev)
            tot_mig += D.allreduce_sum(st["migrated"], dist, dev)
            rounds = max(rounds, st["rounds"])
            kms += D.allreduce_max(st["kernel_ms"], dist, dev)
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    value = tot_steps / tot_t
    kernel_rate = tot_steps / world / (kms * 1e-3)
    achieved = 832 * kernel_rate / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{a.workload}: BASELINE.json configs[4] -- AdvectionRK4_3D, rectilinear {dims['nx']}x{dims['ny']}x{dims['nz']} "
                               f"T={dims['nt']} f32 field DOMAIN-DECOMPOSED into {world} X-slabs (+{halo} halo columns), {n_per_gpu} "
                               f"particles/GPU seeded in the rank's own slab, NCCL all-to-all-v migration; dt={dt:g} s x {nsteps} steps",
                   "timed_region": "upload of the particle shard + initial routing + advect kernels + migration rounds + download "
                                   "(wall clock between barriers, max over ranks)",
                   "migrations_per_pass": tot_mig / a.steps, "advect_rounds_per_pass": rounds,
                   "kernel_ms_per_pass_max_rank": kms / a.steps},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": n * 48 * world, "d2h_bytes_per_step": n * 48 * world,
                "note": "the mode-D pass IS end-to-end: host particle arrays in, host arrays out"},
        "gpu_launches": int(a.steps * rounds * 4),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_particle_step": 832,
                     "note": "per-GPU advect-kernel rate (kernel time only)"},
        "clocks": clk.summary(),
    }  # fmt: skip
    print(json.dumps(line))


# name -> spec.  bytes: algorithmic bytes per particle-step (SURVEY.md 8d / BASELINE.md 4)
WORKLOADS = {
    "c2": dict(field=c2_field, fkw=dict(nx=1440, ny=720, nz=50, nt=3), particles=c2_particles, n=1_000_000, dt=600.0,
               nsteps=144, kernels=["AdvectionRK4_3D"], bytes=832,
               desc="BASELINE.json configs[1] -- AdvectionRK4_3D, rectilinear A-grid 1440x720x50 T=3 f32 U,V,W, spherical"),
    "c2_small": dict(field=c2_field, fkw=dict(nx=360, ny=180, nz=20, nt=3), particles=c2_particles, n=100_000, dt=600.0,
                     nsteps=144, kernels=["AdvectionRK4_3D"], bytes=832, desc="small functional variant of c2"),
    "ns": dict(field=ns_field_device, fkw=dict(nx=4320, ny=2160, nz=50, nt=3), particles=c2_particles, n=10_000_000, dt=600.0,
               nsteps=144, kernels=["AdvectionRK4_3D"], bytes=832, device_field=True,
               roofline_note="real DRAM traffic is 15 % of the algorithmic bytes (per-lane corner cache); the kernel is issue- and "
                             "refill-latency-bound (ncu, profiles/r01d_ncu_summary_ns.txt: issue slots 50 %, fp64 pipe 35 %, L2 hit 13 %)",
               desc="BASELINE.json north_star target -- AdvectionRK4_3D, 1e7 particles on a 1/12 deg rectilinear A-grid "
                    "4320x2160x50 T=3 f32 U,V,W (16.8 GB, generated in HBM), spherical"),
    "c3": dict(field=c3_field, fkw=dict(nx=1442, ny=1021, nt=3), particles=c3_particles, n=10_000_000, dt=3600.0,
               nsteps=48, kernels=["AdvectionRK4"], bytes=320,
               roofline_note="not HBM-bound: 9 double-precision sin/cos + 5 sqrt + ~8 divisions per sample are the reference's own "
                             "arithmetic (ncu, profiles/r01d_ncu_summary_c3_small.txt: DRAM 0.2 % of peak, L1/L2 hit 90 %, cos+sin 26 % of "
                             "the executed instructions, issue slots 40 %, fp64 pipe 23 %)",
               desc="BASELINE.json configs[2] -- AdvectionRK4, curvilinear C-grid ORCA025 shape 1442x1021 T=3, f32 lon/lat, "
                    "CGrid_Velocity + hint/spatial-hash search, spherical"),
    "c3_small": dict(field=c3_field, fkw=dict(nx=362, ny=292, nt=3), particles=c3_particles, n=200_000, dt=3600.0,
                     nsteps=48, kernels=["AdvectionRK4"], bytes=320, desc="small functional variant of c3"),
    "c4": dict(field=c2_field, fkw=dict(nx=1440, ny=720, nz=50, nt=3), particles=c2_particles, n=10_000_000, dt=600.0,
               nsteps=144, kernels=["AdvectionRK4_3D", "DiffusionUniformKh"], bytes=832, kh=(100.0, 50.0),
               desc="BASELINE.json configs[3] -- fused AdvectionRK4_3D + DiffusionUniformKh (Kh 100/50 m2/s) on the config-2 field"),
    "c4_small": dict(field=c2_field, fkw=dict(nx=360, ny=180, nz=20, nt=3), particles=c2_particles, n=100_000, dt=600.0,
                     nsteps=144, kernels=["AdvectionRK4_3D", "DiffusionUniformKh"], bytes=832, kh=(100.0, 50.0),
                     desc="small functional variant of c4"),
}  # fmt: skip


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
def _oracle_fieldset(w, field):
    """Oracle-side field description, built once per process (the curvilinear spatial hash is one-time
    grid setup in the reference too, _core/basegrid.py:192-216, and is kept out of the timed passes)."""
    from oracle import parcels_oracle as po

    if "ofs" not in _G:
        pad = field.get("padding", ("low", "low", "high"))
        g = po.OGrid(field["lon"], field["lat"], field["depth"], mesh=field["mesh"], offsets=tuple(int(p == "low") for p in pad))
        consts = {"Kh_zonal": w["kh"][0], "Kh_meridional": w["kh"][1]} if "kh" in w else None
        _G["ofs"] = po.OFieldSet(g, field["U"], field["V"], field["W"], time=field["times"], interp=field.get("interp", "linear"),
                                 constants=consts)  # fmt: skip
        if g.curvilinear:
            from oracle import curvilinear_oracle as co

            co.get_hash(g)
    return _G["ofs"]


def _oracle_pass(w, field, parts):
    from oracle import parcels_oracle as po

    fs = _oracle_fieldset(w, field)
    pd = po.create_particle_data(parts["x"], parts["y"], parts["z"], parts["t"], ngrids=fs.ngrids)
    kmap = {"AdvectionRK4_3D": po.AdvectionRK4_3D, "AdvectionRK4": po.AdvectionRK4}
    kern = [po.DiffusionUniformKh() if k == "DiffusionUniformKh" else kmap[k] for k in w["kernels"]] + [po.DeleteOnError]
    t0 = time.perf_counter()
    steps = po.pset_execute(pd, fs, kern, w["dt"], runtime=w["dt"] * w["nsteps"])
    return steps, time.perf_counter() - t0


def _c_port_pass(w, field, parts):
    """strong CPU baseline: the C + OpenMP restatement (oracle/advect_rk4_3d.c) on all host cores"""
    from oracle import c_port
    from oracle import parcels_oracle as po

    pd = po.create_particle_data(parts["x"], parts["y"], parts["z"], parts["t"])
    g = po.OGrid(field["lon"], field["lat"], field["depth"], mesh=field["mesh"])
    c_port.advect_rk4_3d(lon=field["lon"], lat=field["lat"], depth=field["depth"], time=field["times"], U=field["U"], V=field["V"],
                         W=field["W"], spherical=g.spherical, deg2m=g.deg2m, pdata=dict(pd), dt=w["dt"], endtime=w["dt"])  # warm-up: 1 step
    pd = po.create_particle_data(parts["x"], parts["y"], parts["z"], parts["t"])
    t0 = time.perf_counter()
    steps = c_port.advect_rk4_3d(lon=field["lon"], lat=field["lat"], depth=field["depth"], time=field["times"], U=field["U"],
                                 V=field["V"], W=field["W"], spherical=g.spherical, deg2m=g.deg2m, pdata=pd, dt=w["dt"],
                                 endtime=w["dt"] * w["nsteps"])  # fmt: skip
    return steps, time.perf_counter() - t0


def _worker(args):
    seed, n = args
    w, field = _G["w"], _G["field"]
    steps, _ = _oracle_pass(w, field, w["particles"](field, n, seed))
    return steps


def run_reference_arm(a, w, field):
    """Reference arm: the reference's algorithm (oracle port) on all host cores (rank 0 only)."""
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    _G["w"], _G["field"] = w, field
    _oracle_fieldset(w, field)  # incl. the spatial hash: built once, before forking
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def one_pass(k, per_=None):
            t0 = time.perf_counter()
            steps = sum(pool.map(_worker, [(1000 * k + c, per_ or per) for c in range(cores)]))
            return steps, time.perf_counter() - t0

        # size the per-step sample so that one bench step takes ~a.ref_step_seconds on THIS box (the whole
        # --steps K --warmup W run must end within a few minutes): calibrate on a small pass first
        per = a.ref_particles_per_core
        if per <= 0:
            s0, t0_ = one_pass(999, 200)
            per = int(np.clip(a.ref_step_seconds * (s0 / t0_) / (cores * w["nsteps"]), 50, 20000))

        for k in range(a.warmup):
            one_pass(k)
        tot_steps, tot_t = 0, 0.0
        for k in range(a.steps):
            s, t = one_pass(100 + k)
            tot_steps += s
            tot_t += t
    v = tot_steps / tot_t
    sample = f"{cores} procs x {per} particles x {w['nsteps']} dt-steps per bench step (same field, dt, kernels)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "particle-steps/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / max(a.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{a.workload}: {w['desc']}", "kernels": w["kernels"],
                   "note": "oracle port of the reference's NumPy path (the reference is pure Python and absent on the GPU box)"},
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
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS) + ["c5", "c5_small"])
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="particles of the cpu_baseline sample")
    ap.add_argument("--ref-particles-per-core", type=int, default=0, help="0: calibrate to --ref-step-seconds per bench step")
    ap.add_argument("--ref-step-seconds", type=float, default=5.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", type=int, default=0,
                    help="experiment: the end-to-end arm runs pb_advect_host with this many pipelined chunks and copies the "
                         "result back with the same call (ParticleSet.pipeline_chunks, eager_host)")
    ap.add_argument("--sorted", action="store_true",
                    help="experiment: release the particles ordered by grid cell (z, y, x) instead of randomly -- measures what "
                         "spatial coherence between the lanes of a warp is worth (DESIGN.md 4, 'why not Morton-sort')")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload.startswith("c5"):
        if a.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "config 5 has no single-process CPU equivalent; use --workload c2"}))
            return
        if world < 2:
            raise SystemExit("workload c5 is the domain-decomposed mode: launch with torchrun on >= 2 GPUs")
        return run_decomposed_bench(a, rank, local_rank, world)
    w = WORKLOADS[a.workload]
    n_per_gpu = a.particles or w["n"]
    dt, nsteps = w["dt"], w["nsteps"]

    if a.impl == "reference":
        if rank == 0:
            if w.get("device_field"):
                print(json.dumps({"impl": "reference", "unavailable": f"workload {a.workload} builds its field in HBM; the CPU arm is "
                                                                       "timed on --workload c2 (same kernels, same arithmetic)"}))  # fmt: skip
            else:
                run_reference_arm(a, w, w["field"](**w["fkw"]))
        return

    import torch

    import parcels_b200 as pb
    from parcels_b200 import build
    from parcels_b200.kernels import SCHEMES

    build.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    if w.get("device_field"):
        torch.cuda.set_device(local_rank)
        field = w["field"](device=local_rank, **w["fkw"])
        a.no_cpu_baseline = True  # the NumPy / C arms need the field on the host: they are timed on c2
    else:
        field = w["field"](**w["fkw"])
    fs = pb.FieldSet.from_arrays(lon=field["lon"], lat=field["lat"], depth=field["depth"], time=field["times"],
                                 U=field["U"], V=field["V"], W=field["W"], mesh=field["mesh"],
                                 interp_method=field.get("interp", "linear"),
                                 padding=field.get("padding", ("low", "low", "high")))  # fmt: skip
    diffusion = "DiffusionUniformKh" in w["kernels"]
    if diffusion:
        fs.add_constant_field("Kh_zonal", w["kh"][0], mesh=field["mesh"])
        fs.add_constant_field("Kh_meridional", w["kh"][1], mesh=field["mesh"])
    parts = w["particles"](field, n_per_gpu, 1 + rank)  # weak scaling: every rank owns its own shard
    if a.sorted and np.ndim(field["lon"]) == 1:
        cell = np.searchsorted(field["lon"], parts["x"]).astype(np.int64)
        cell += len(field["lon"]) * np.searchsorted(field["lat"], parts["y"])
        if field["depth"] is not None:
            cell += len(field["lon"]) * len(field["lat"]) * np.searchsorted(field["depth"], parts["z"])
        order = np.argsort(cell, kind="stable")
        parts = {k: v[order] for k, v in parts.items()}
    ps = pb.ParticleSet(fs, x=parts["x"], y=parts["y"], z=parts["z"], t=parts["t"], device=local_rank, seed=1234)
    init = {k: v.copy() for k, v in ps._data.items()}
    eng = fs.engine(local_rank)
    runtime = dt * nsteps
    kernels = [getattr(pb, k) for k in w["kernels"]] + [pb.DeleteParticle]

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    # ---- device-resident arm: particles + fields in HBM; per step: restore snapshot + ONE kernel ----
    ei_last = np.ascontiguousarray(init["ei"][:, -1])
    eng.upload_particles(init, ei_last)
    eng.snapshot()
    plan = pb.particleset.KernelPlan(kernels, fs)
    args = eng.make_args(SCHEMES[w["kernels"][0]], dt, runtime, diffusion=diffusion, delete_on_error=True, kh=plan.kh,
                         kh_spherical=plan.kh_spherical, kh_deg2m=plan.kh_deg2m, seed=1234, rng_call=1,
                         hint_all_zero=fs.grid.curvilinear)  # fmt: skip

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
            rep = eng.last_report()  # waits for this step's kernel (the report is read back every step)
            ksum += rep["kernel_ms"]
            psteps += rep["particle_steps"]
        dev_ms = eng.timer_end_ms()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = reduce(dev_ms, "MAX")
    total_steps = reduce(psteps, "SUM")
    value = total_steps / (dev_ms * 1e-3)
    kernel_ms = ksum / a.steps
    steps_per_launch = psteps / a.steps

    # ---- end-to-end arm: public API with host arrays; H2D + D2H of the particle SoA every step ----
    e2e = None
    if not a.no_e2e:
        k_e2e = min(a.steps, 5) if n_per_gpu > 2_000_000 else a.steps
        def pinned(v):  # page-locked host copy (the contract's "inputs from pinned host memory")
            return torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy()

        fresh = [{k: pinned(v) for k, v in init.items()} for _ in range(k_e2e)]  # host input batches, made before timing
        if a.pipeline > 1:
            ps.pipeline_chunks, ps.eager_host = a.pipeline, True
        for _ in range(min(a.warmup, 2)):
            ps._data = {k: v.copy() for k, v in init.items()}
            ps.execute(kernels, dt=dt, runtime=runtime)
            _ = ps._data
        barrier()
        e2e_steps = 0
        t0 = time.perf_counter()
        for i in range(k_e2e):
            ps._data = fresh[i]
            ps.execute(kernels, dt=dt, runtime=runtime)
            result = ps._data  # device -> host read of the step's result (the particle SoA after the pass)
            assert not ps._host_stale and len(result["x"]) > 0
            e2e_steps += ps.last_report["particle_steps"]
        barrier()
        e2e_s = reduce(time.perf_counter() - t0, "MAX")
        n = n_per_gpu
        e2e = {"value": reduce(e2e_steps, "SUM") / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": n * (6 * 4 + 8 + 4 + 4 + 8),
               "d2h_bytes_per_step": n * (6 * 4 + 8 + 4 + 4), "steps": k_e2e,
               "api": f"parcels_b200.ParticleSet.execute([{', '.join(w['kernels'])}, DeleteParticle], dt={dt:g}, runtime={runtime:g})"}  # fmt: skip
        if a.pipeline > 1:
            e2e["pipeline_chunks"] = a.pipeline

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    achieved = w["bytes"] * steps_per_launch / (kernel_ms * 1e-3) / 1e9
    fbytes = sum(int(np.prod(field[k].shape)) * 4 for k in ("U", "V", "W") if field.get(k) is not None)
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{a.workload}: {w['desc']}; {n_per_gpu} particles/GPU, dt={dt:g} s x {nsteps} dt-steps per pass"
                        + ("; particles released in grid-cell order (--sorted experiment)" if a.sorted else ""),
            "kernels": w["kernels"] + ["DeleteParticle"], "particles_per_gpu": n_per_gpu, "dt_steps_per_pass": nsteps,
            "l2_policy": f"inputs larger than L2 ({fbytes / 1e9:.2f} GB field, {52 * n_per_gpu / 1e6:.0f} MB particle SoA; particles "
                         "re-seeded from an HBM snapshot every pass)",
            "wall_ms_per_step": wall_ms / a.steps, "kernel_ms_per_launch": kernel_ms,
            "corner_cache_refills_per_launch": rep["cache_refills"], "deleted_per_launch": rep["n_deleted"],
        },
        "e2e": e2e,
        "gpu_launches": a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": dram_traffic_per_launch(a.workload), "peak_source": peak_src,
                     "algorithmic_bytes_per_particle_step": w["bytes"], "particle_steps_per_launch": steps_per_launch,
                     **({"note": w["roofline_note"]} if "roofline_note" in w else {})},
        "clocks": clk.summary(),
    }  # fmt: skip
    if not a.no_cpu_baseline:
        sample = w["particles"](field, a.cpu_sample, 1)
        _oracle_fieldset(w, field)
        s, t = _oracle_pass(w, field, sample)
        line["cpu_baseline"] = {"value": s / t, "unit": "particle-steps/s", "cores": 1, "kind": "port",
                                "sample": f"{a.cpu_sample} particles of the same workload, {nsteps} dt-steps, NumPy oracle port of "
                                          f"the reference path ({t:.1f} s)"}  # fmt: skip
        if w["kernels"] == ["AdvectionRK4_3D"] and field.get("interp", "linear") == "linear":
            cores = len(os.sched_getaffinity(0))
            big = w["particles"](field, min(n_per_gpu, 400_000), 1)
            s2, t2 = _c_port_pass(w, field, big)
            line["cpu_baseline_strong"] = {"value": s2 / t2, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                                           "sample": f"{len(big['x'])} particles, {nsteps} dt-steps, C + OpenMP restatement "
                                                     f"(oracle/advect_rk4_3d.c, same arithmetic, all host cores; {t2:.1f} s)"}  # fmt: skip
    print(json.dumps(line))


if __name__ == "__main__":
    main()
