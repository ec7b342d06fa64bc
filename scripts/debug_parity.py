"""Debug helper: per-case diff summary engine vs oracle (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from engine_run import run_engine, ulp_diff_f32
from oracle_run import load_case, run_oracle

for name in sys.argv[1:]:
    c = load_case(name)
    ps, err = run_engine(c)
    pd, oerr = run_oracle(c)
    d = ps._data
    print(f"== {name}: engine n={len(d['x'])} err={err!r}; oracle n={len(pd['x'])} err={oerr}")
    ids_e, ids_o = set(d["particle_id"].tolist()), set(pd["particle_id"].tolist())
    print("   only engine:", sorted(ids_e - ids_o)[:10], " only oracle:", sorted(ids_o - ids_e)[:10])
    common = sorted(ids_e & ids_o)
    ie = np.searchsorted(d["particle_id"], common); io = np.searchsorted(pd["particle_id"], common)
    for key in ("state", "t", "ei"):
        a, b = d[key][ie], pd[key][io]
        bad = np.where((a != b).reshape(len(common), -1).any(axis=1))[0]
        print(f"   {key}: {len(bad)} mismatches", [(common[k], a[k].tolist(), b[k].tolist()) for k in bad[:5]])
    for key in "xyz":
        u = ulp_diff_f32(d[key][ie], pd[key][io])
        print(f"   {key}: max ulp {u.max()}  hist", np.bincount(np.minimum(u, 10))[:11].tolist())
    print("   report", ps.last_report)
