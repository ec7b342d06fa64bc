import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from engine_run import run_engine, ulp_diff_f32
from oracle_run import load_case, run_oracle
name = sys.argv[1]
for nsteps in (1, 2, 3):
    c = load_case(name)
    c["segments"] = [dict(runtime=c["dt"] * nsteps)]
    c["delete_on_error"] = False
    ps, err = run_engine(c)
    pd, oerr = run_oracle(c)
    d = ps._data
    print(f"== steps={nsteps}: engine err={err!r} oracle err={oerr}")
    for key in ("state", "ei", "x", "y", "t"):
        a, b = d[key].reshape(len(d["x"]), -1), pd[key].reshape(len(d["x"]), -1)
        bad = np.where((a != b).any(axis=1))[0]
        print(f"   {key}: {len(bad)} mismatches", [(int(k), a[k].tolist(), b[k].tolist()) for k in bad[:6]])
    if nsteps == 1:
        bad = np.where(d["x"] != pd["x"])[0][:5]
        for k in bad:
            print("   particle", k, "start", c["x"][k], c["y"][k], "engine", d["x"][k], d["y"][k], d["ei"][k], d["state"][k], "oracle", pd["x"][k], pd["y"][k], pd["ei"][k], pd["state"][k])
