"""Helper of tests/test_install_cpu.py (run in a subprocess whose PB_LIB points at the host simulation of the kernel sources, or
on a GPU box): the reference's OWN ``ParticleSet.execute`` (under the stub harness, oracle/ref_harness.py) runs the same inputs
twice -- untouched, and with ``parcels_b200.install()`` patching ``Kernel.execute`` -- and the two ``pset._data`` are compared."""

import json
import sys
import warnings

import numpy as np

from oracle import ref_harness as rh
from oracle_run import load_case

rh.install()
import parcels._core.kernel as rk  # noqa: E402
import parcels._core.statuscodes as rcodes  # noqa: E402

import parcels_b200 as pb  # noqa: E402

K = rh.kernels()
out = {}


def fieldset(c, **kw):
    return rh.build_fieldset(lon=c["lon"], lat=c["lat"], depth=c["depth"], times=c["times"], U=c["U"], V=c["V"], W=c["W"], mesh=c["mesh"],
                             constants=c.get("constants"), **kw)  # fmt: skip


def run(c, kernels, patched, dt=None, **exec_kw):
    (pb.install if patched else pb.uninstall)()
    fs = fieldset(c)
    ps = rh.make_pset(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    err = ""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ps.execute(kernels, dt=np.timedelta64(int(dt or c["dt"]), "s"), verbose_progress=False, **exec_kw)
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__module__}.{type(e).__name__}"
    pb.uninstall()
    return ps, err


def same(a, b, ulp=0):
    from engine_run import ulp_diff_f32

    res = {}
    for k in ("particle_id", "state", "t", "ei", "dt"):
        res[k] = bool(a._data[k].shape == b._data[k].shape and np.array_equal(a._data[k], b._data[k]))
    for k in "xyz":
        ok = a._data[k].shape == b._data[k].shape
        res[k] = bool(ok and (len(a._data[k]) == 0 or ulp_diff_f32(a._data[k], b._data[k]).max() <= ulp))
    return res


def user_delete(particles, fieldset):  # the reference's own idiom (tests/common_kernels.py:12-13): a USER kernel, not a token
    particles[particles.state >= 50].state = rcodes.StatusCode.Delete


# (untouched reference runs use the user handler -- the engine's DeleteParticle TOKEN is the same handler fused into the launch)
# 1. built-ins only, flat mesh, 3-D, f32 coordinates / f64 data (the v3-golden dtype combination): bit-exact
c = load_case("flat_f32c_f64d")
a, ea = run(c, [K.AdvectionRK4_3D, user_delete], False, runtime=np.timedelta64(150, "s"))
b, eb = run(c, [K.AdvectionRK4_3D, pb.DeleteParticle], True, runtime=np.timedelta64(150, "s"))
out["builtins_flat"] = dict(same=same(a, b), err=[ea, eb], n=[len(a._data["x"]), len(b._data["x"])], patched=bool(b.__dict__.get("_b200_pset") is not None))

# 2. the reference's own kernel list with a USER error handler: host loop control + device built-ins
a, ea = run(c, [K.AdvectionRK4_3D, user_delete], False, runtime=np.timedelta64(150, "s"))
b, eb = run(c, [K.AdvectionRK4_3D, user_delete], True, runtime=np.timedelta64(150, "s"))
out["user_handler"] = dict(same=same(a, b), err=[ea, eb], n=[len(a._data["x"]), len(b._data["x"])])

# 3. no handler: the reference raises FieldOutOfBoundError (its OWN class) at the same iteration, particles left where it leaves them
a, ea = run(c, [K.AdvectionRK4_3D], False, runtime=np.timedelta64(150, "s"))
b, eb = run(c, [K.AdvectionRK4_3D], True, runtime=np.timedelta64(150, "s"))
out["raises"] = dict(same=same(a, b), err=[ea, eb])

# 4. spherical mesh, f64 coordinates / f32 data (config-2 dtypes; the specialised RK4 kernel): <= 2 float32 ulp (cos)
c = load_case("c2_small")
a, ea = run(c, [K.AdvectionRK4_3D, user_delete], False, runtime=np.timedelta64(7200, "s"))
b, eb = run(c, [K.AdvectionRK4_3D, pb.DeleteParticle], True, runtime=np.timedelta64(7200, "s"))
out["c2_small"] = dict(same=same(a, b, ulp=2), err=[ea, eb], n=[len(a._data["x"]), len(b._data["x"])],
                       variant=b.__dict__["_b200_pset"].last_report.get("kernel_variant"))

# 5. several output intervals through the reference's own outer loop (a duck output file counting the writes)
class Recorder:
    outputdt = 50.0

    def __init__(self):
        self.rows = []

    def write(self, pset, t, indices=None):
        self.rows.append((float(t), pset._data["x"].copy()))

    def set_metadata(self, *a, **k):
        pass

    metadata = {}
    path = "memory"

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


c = load_case("flat_f32c_f64d")
ra, rb = Recorder(), Recorder()
a, ea = run(c, [K.AdvectionRK4_3D, user_delete], False, runtime=np.timedelta64(150, "s"), output_file=ra)
b, eb = run(c, [K.AdvectionRK4_3D, pb.DeleteParticle], True, runtime=np.timedelta64(150, "s"), output_file=rb)
out["output_intervals"] = dict(same=same(a, b), writes=[len(ra.rows), len(rb.rows)],
                               rows_equal=bool(len(ra.rows) == len(rb.rows) and all(t1 == t2 and np.array_equal(x1, x2) for (t1, x1), (t2, x2) in zip(ra.rows, rb.rows))))

# 6. restored: after uninstall the reference's own method is back
out["uninstalled"] = not getattr(rk.Kernel.execute, "_b200_patched", False)
print("INSTALL_RESULT " + json.dumps(out))
