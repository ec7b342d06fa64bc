import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from engine_run import make_fieldset, run_engine
from oracle import parcels_oracle as po
from oracle_run import load_case, oracle_fieldset
name = sys.argv[1]
c = load_case(name)
fs = make_fieldset(c); ofs = oracle_fieldset(c); eng = fs.engine(0)
n = len(c["x"]); dt = c["dt"]
x = c["x"].astype(np.float32); y = c["y"].astype(np.float32); z = c["z"].astype(np.float32)
pd = po.create_particle_data(x, y, z, 0.0)
view = po.View(pd, np.ones(n, bool))
t0 = np.zeros(n)
def both(t, zz, yy, xx, f32, hint, no_hint):
    u, v, w, ei, st = eng.sample_velocity(t, zz, yy, xx, three_d=False, positions_are_f32=f32, ei_hint=hint, no_hint=no_hint)
    (ou, ov) = po.eval_uvw(ofs, t, zz, yy, xx, view, False)
    oei = pd["ei"][:, -1].copy()
    sc = np.nanmax(np.abs(ou[np.isfinite(ou)]))
    bad = np.where((ei != oei) | ~(np.abs(u - ou) <= 1e-6 * sc) | ~(np.abs(v - ov) <= 1e-6 * sc))[0]
    print("   eval: mismatches", len(bad), "nonfinite oracle", (~np.isfinite(ou)).sum(), [(int(k), ei[k], oei[k], u[k], ou[k]) for k in bad[:4]])
    print("      particle 5: dev", u[5], v[5], ei[5], "oracle", ou[5], ov[5], oei[5], "pos", xx[5], yy[5], t[5])
    return (u, v, ei), (ou, ov, oei)
(u1, v1, e1), (ou1, ov1, oe1) = both(t0, z, y, x, True, np.zeros(n, np.int32), True)
x1 = x + u1 * 0.5 * dt; y1 = y + v1 * 0.5 * dt
(u2, v2, e2), _ = both(t0 + 0.5 * dt, z, y1, x1, False, e1, False)
x2 = x + u2 * 0.5 * dt; y2 = y + v2 * 0.5 * dt
(u3, v3, e3), _ = both(t0 + 0.5 * dt, z, y2, x2, False, e2, False)
x3 = x + u3 * dt; y3 = y + v3 * dt
(u4, v4, e4), _ = both(t0 + dt, z, y3, x3, False, e3, False)
xn = (x + ((u1 + 2 * u2 + 2 * u3 + u4) / 6.0 * dt).astype(np.float32)).astype(np.float32)
c["segments"] = [dict(runtime=dt)]; c["delete_on_error"] = False
ps, err = run_engine(c)
d = ps._data
bad = np.where(d["x"] != xn)[0]
print("kernel vs python-driven RK4 (device evals): mismatches", len(bad), [(int(k), d["x"][k], xn[k], d["ei"][k], e4[k]) for k in bad[:8]])
print("cells per stage for first bad:", [(int(k), e1[k], e2[k], e3[k], e4[k]) for k in bad[:8]])
