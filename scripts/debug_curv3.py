import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from engine_run import make_fieldset, run_engine
from oracle_run import load_case, run_oracle
name = sys.argv[1]
res = []
for rep in range(3):
    c = load_case(name)
    c["segments"] = [dict(runtime=c["dt"])]; c["delete_on_error"] = False
    ps, err = run_engine(c)
    res.append(ps._data["x"].copy())
pd, oerr = run_oracle(c)
print("run0 vs run1 mismatches", (res[0] != res[1]).sum(), "run1 vs run2", (res[1] != res[2]).sum())
for r in res: print("  vs oracle", (r != pd["x"]).sum(), "particle5", r[5], pd["x"][5])
