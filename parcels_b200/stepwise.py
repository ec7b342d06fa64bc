"""``Kernel.execute`` for kernel lists that mix built-in kernels with USER Python kernels
(the common idiom of the reference's own tests, e.g. ``[AdvectionRK4_3D, DeleteParticle]`` with a user-defined
handler, periodic-boundary or ageing kernels; reference ``_core/kernel.py:174-247``).

The loop control of one ``Kernel.execute`` call runs here, step by step, exactly as in the reference; every
BUILT-IN kernel of the list is still executed on the device (``pb_advect(kernels_only=1)``: one iteration's kernel
functions on the evaluated particles), user functions are called on the host with a ``ParticleSetView``.  The
price is one host<->device round trip of the particle SoA per time step; lists made of built-ins only never come
here (they run the whole loop in one kernel launch).
"""

from __future__ import annotations

import numpy as np

from .particlesetview import ParticleSetView
from .statuscodes import ERRORS_TO_THROW, StatusCode, raise_for_state


def kernel_execute_stepwise(pset, plan, endtime: float, dt: float):
    d = pset._data
    fs = pset.fieldset
    eng = fs.engine(pset.device)
    eng.claim(pset)
    sign = 1 if dt > 0 else -1
    d["state"][:] = StatusCode.Evaluate
    steps = 0
    # RK45 mode (reference kernel.py:118-120,224-226: `hasattr(fieldset, "RK45_tol")`): dt <- next_dt after the position update,
    # and dt is NOT reset to the nominal step at the end of the iteration
    rk45_mode = "RK45_tol" in fs.context
    while len(d["x"]) > 0 and np.any(np.isin(d["state"], [StatusCode.Evaluate, StatusCode.Repeat])):
        tte = sign * (endtime - d["t"])
        evaluate = np.isin(d["state"], [StatusCode.Success, StatusCode.Evaluate]) & (tte >= 0)
        if not np.any(evaluate):
            break
        # adapt dt to end exactly on endtime (kernel.py:199-203)
        d["dt"][:] = np.maximum(np.minimum(d["dt"], tte), 0) if sign == 1 else np.minimum(np.maximum(d["dt"], -tte), 0)
        steps += int(np.count_nonzero(evaluate))
        for item in plan.items:
            if item[0] in ("device", "advdiff"):
                _device_kernels(pset, eng, item, dt, endtime)
            elif item[0] == "rk45":
                _device_rk45(pset, eng, item[1], dt, endtime)
            else:
                f = item[1]
                f(ParticleSetView(d, evaluate, fs), fs)
                repeat = d["state"] == StatusCode.Repeat
                while np.any(repeat):  # kernel.py:213-216
                    f(ParticleSetView(d, repeat, fs), fs)
                    repeat = d["state"] == StatusCode.Repeat
        # position update only for particles still in a normal state (kernel.py:108-116,220-222)
        upd = evaluate & np.isin(d["state"], [StatusCode.Evaluate, StatusCode.Success])
        if np.any(upd):
            d["x"][upd] += d["dx"][upd]
            d["y"][upd] += d["dy"][upd]
            d["z"][upd] += d["dz"][upd]
            d["t"][upd] += d["dt"][upd]
            d["dx"][upd] = 0
            d["dy"][upd] = 0
            d["dz"][upd] = 0
            if rk45_mode:
                d["dt"][upd] = d["next_dt"][upd]
        if not rk45_mode:
            d["dt"][:] = dt
        d["state"][(d["state"] == StatusCode.Evaluate) & (d["t"] == endtime)] = StatusCode.EndofLoop
        dele = np.where(d["state"] == StatusCode.Delete)[0]
        if len(dele) > 0:
            pset.remove_indices(dele)
            d = pset._data
        if np.any(d["state"] == StatusCode.StopAllExecution):
            break
        for code in ERRORS_TO_THROW:
            hit = d["state"] == code
            if np.any(hit):
                raise_for_state(code, d["z"][hit], d["y"][hit], d["x"][hit], d["t"][hit])
    pset.last_report = dict(particle_steps=steps, mode="stepwise")


def _device_kernels(pset, eng, item, dt, endtime):
    """One iteration of the built-in kernels ``item`` on the device: the kernel applies the reference's own
    evaluate mask (state in {Success, Evaluate} and time-to-endtime >= 0) and dt clamp, accumulates dx/dy/dz and
    updates state/ei; nothing else."""
    d = pset._data
    pset._rng_call += 1
    ei_last = np.ascontiguousarray(d["ei"][:, -1])
    hint_all_zero = False
    g = pset.fieldset.grid
    if g.curvilinear:
        from .particleset import _hint_all_zero

        sign = 1 if dt > 0 else -1
        hint_all_zero = _hint_all_zero(
            ei_last, lambda s_: np.isin(d["state"][s_], [StatusCode.Success, StatusCode.Evaluate]) & (sign * (endtime - d["t"][s_]) >= 0), g.xdim
        )
    eng.upload_particles(d, ei_last)
    from .particleset import _batch_levels

    sign_ = 1 if dt > 0 else -1
    two_levels = _batch_levels(
        pset.fieldset, d, lambda: np.isin(d["state"], [StatusCode.Success, StatusCode.Evaluate]) & (sign_ * (endtime - d["t"]) >= 0)
    )
    if item[0] == "advdiff":  # AdvectionDiffusionM1 / EM (pb_advect_diffusion)
        args = eng.make_advdiff_args(dt=dt, endtime=endtime, seed=pset.seed, rng_call=pset._rng_call, resume=True, kernels_only=True,
                                     batch_levels=two_levels, **item[1])  # fmt: skip
    else:
        _, scheme, diffusion, plan = item
        args = eng.make_args(scheme, dt, endtime, diffusion=diffusion, kh=plan.kh, kh_spherical=plan.kh_spherical,
                             kh_deg2m=plan.kh_deg2m, seed=pset.seed, rng_call=pset._rng_call, hint_all_zero=hint_all_zero,
                             resume=True, kernels_only=True, batch_levels=two_levels)  # fmt: skip
    rep = eng.advect(args)
    eng.download_particles(d, ei_last)
    d["ei"][:, -1] = ei_last
    if rep["n_out_of_time"] > 0:  # the reference flags the whole evaluated view (index_search.py:85-86, field.py:31-44)
        sign = 1 if dt > 0 else -1
        view = np.isin(d["state"], [StatusCode.Success, StatusCode.Evaluate, StatusCode.ErrorOutsideTimeInterval]) & (
            sign * (endtime - d["t"]) >= 0
        )
        d["state"][view] = StatusCode.ErrorOutsideTimeInterval


def _device_rk45(pset, eng, params, dt, endtime):
    """One iteration of AdvectionRK45 on the device (``pb_advect_rk45`` with kernels_only): every evaluated particle's attempts until
    its step is accepted -- the reference's `while state == Repeat` re-runs of the kernel (kernel.py:212-216) -- with dx / dy,
    dt, next_dt, state and ei written back; the host does the position update and `dt <- next_dt`."""
    d = pset._data
    ei_last = np.ascontiguousarray(d["ei"][:, -1])
    eng.upload_particles(d, ei_last)
    dt_arr = np.ascontiguousarray(d["dt"], dtype=np.float64)
    ndt_arr = np.ascontiguousarray(d["next_dt"], dtype=np.float64)
    tol, min_dt, max_dt = params
    hint_all_zero = False
    g = pset.fieldset.grid
    if g.curvilinear:
        from .particleset import _hint_all_zero

        sign = 1 if dt > 0 else -1
        hint_all_zero = _hint_all_zero(
            ei_last, lambda s_: np.isin(d["state"][s_], [StatusCode.Success, StatusCode.Evaluate]) & (sign * (endtime - d["t"][s_]) >= 0), g.xdim
        )
    from .particleset import _batch_levels

    sign_ = 1 if dt > 0 else -1
    two_levels = _batch_levels(
        pset.fieldset, d, lambda: np.isin(d["state"], [StatusCode.Success, StatusCode.Evaluate]) & (sign_ * (endtime - d["t"]) >= 0)
    )
    rep = eng.advect_rk45(dt, endtime, tol, min_dt, max_dt, dt_arr, ndt_arr, next_dt_is_f32=d["next_dt"].dtype == np.float32,
                          kernels_only=True, resume=True, hint_all_zero=hint_all_zero, batch_levels=two_levels)  # fmt: skip
    eng.download_particles(d, ei_last)
    d["ei"][:, -1] = ei_last
    d["dt"][:] = dt_arr
    d["next_dt"][:] = ndt_arr
    if rep["n_error"] > 0:
        stuck = np.where(d["state"] == StatusCode.Error)[0]
        raise RuntimeError(f"AdvectionRK45: particles {stuck[:10]} have dt == 0 before endtime={endtime}; reset pset.dt before continuing")
