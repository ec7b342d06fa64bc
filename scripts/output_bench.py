"""Output-path measurement: one ParticleSet.execute with a Parquet ParticleFile, output every
`--every` dt-steps, device-resident intervals (rows selected + compacted on the GPU) vs the host path (full SoA
download every interval, selection with NumPy).  Prints one JSON line.  Not a bench.py metric."""

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
import parcels_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--every", type=int, default=6)
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    w = bench.WORKLOADS["c2_small" if a.small else "c2"]
    field = w["field"](**w["fkw"])
    fs = pb.FieldSet.from_arrays(lon=field["lon"], lat=field["lat"], depth=field["depth"], time=field["times"], U=field["U"],
                                 V=field["V"], W=field["W"], mesh=field["mesh"])  # fmt: skip
    parts = w["particles"](field, a.n, 1)
    dt = w["dt"]
    out = {"particles": a.n, "dt_steps": a.steps, "output_every": a.every, "writes": a.steps // a.every + 1}
    tmp = tempfile.mkdtemp()
    for mode in ("host", "device", "device_null", "host_null"):
        # the host path is forced by an extra (unwritten) particle variable: extra variables live in host arrays
        pclass = pb.Particle if mode.startswith("device") else pb.Particle.add_variable(pb.Variable("tmp", np.float32, to_write=False))
        best = None
        for rep in range(3):
            ps = pb.ParticleSet(fs, pclass=pclass, x=parts["x"], y=parts["y"], z=parts["z"], t=parts["t"])
            if mode.endswith("null"):  # selection + D2H only, no Parquet encoding

                class Null:
                    outputdt = a.every * dt
                    rows = 0

                    def write(self, pset, t):
                        cols, _ = pset._output_columns(float(t), ["t", "z", "y", "x", "particle_id"])
                        self.rows += len(cols["t"])

                pf = Null()
            else:
                pf = pb.ParticleFile(os.path.join(tmp, f"{mode}{rep}.parquet"), outputdt=a.every * dt, compression=None)
            fs.engine(0).synchronize()
            t0 = time.perf_counter()
            ps.execute([pb.AdvectionRK4_3D, pb.DeleteParticle], dt=dt, runtime=a.steps * dt, output_file=pf)
            n_end = len(ps._data["x"])  # final state on the host in both modes
            el = time.perf_counter() - t0
            best = el if best is None else min(best, el)
        out[mode] = {"seconds": best, "rows": pf.rows_written if hasattr(pf, "rows_written") else pf.rows, "n_end": n_end}
    out["speedup_with_parquet"] = out["host"]["seconds"] / out["device"]["seconds"]
    out["speedup_selection_and_copies"] = out["host_null"]["seconds"] / out["device_null"]["seconds"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
