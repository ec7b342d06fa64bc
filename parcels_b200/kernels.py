"""Built-in kernels, by the reference's names (``parcels.kernels``).

On this engine a built-in kernel is a *token*: ``ParticleSet.execute`` recognises the function
(by identity or, for the reference's own function objects, by ``__name__``), and the whole
kernel list is lowered to ONE fused CUDA kernel that runs the per-particle time loop on the GPU
(``csrc/engine.cu``).  The bodies below never run -- there is no CPU path in this package.
"""

from __future__ import annotations

__all__ = [
    "AdvectionEE",
    "AdvectionRK2",
    "AdvectionRK2_3D",
    "AdvectionRK4",
    "AdvectionRK4_3D",
    "AdvectionRK45",
    "AdvectionDiffusionM1",
    "AdvectionDiffusionEM",
    "DeleteParticle",
    "DiffusionUniformKh",
]


def _device_only(name):
    raise RuntimeError(
        f"{name} is a device kernel token: pass it to ParticleSet.execute(); parcels_b200 has no CPU implementation."
    )


def AdvectionEE(particles, fieldset):  # reference kernels/_advection.py:78-82
    """Explicit Euler advection of (x, y) with fieldset.UV."""
    _device_only("AdvectionEE")


def AdvectionRK2(particles, fieldset):  # reference kernels/_advection.py:20-27
    """Second-order Runge-Kutta advection of (x, y) with fieldset.UV."""
    _device_only("AdvectionRK2")


def AdvectionRK2_3D(particles, fieldset):  # reference kernels/_advection.py:30-39
    """Second-order Runge-Kutta advection of (x, y, z) with fieldset.UVW."""
    _device_only("AdvectionRK2_3D")


def AdvectionRK4(particles, fieldset):  # reference kernels/_advection.py:42-55
    """Fourth-order Runge-Kutta advection of (x, y) with fieldset.UV."""
    _device_only("AdvectionRK4")


def AdvectionRK4_3D(particles, fieldset):  # reference kernels/_advection.py:58-75
    """Fourth-order Runge-Kutta advection of (x, y, z) with fieldset.UVW."""
    _device_only("AdvectionRK4_3D")


def AdvectionRK45(particles, fieldset):  # reference kernels/_advection.py:85-155
    """Adaptive Runge-Kutta-Fehlberg 4(5) advection of (x, y) with fieldset.UV; needs a ``next_dt`` particle Variable and the
    fieldset context RK45_tol (m), RK45_min_dt, RK45_max_dt (s)."""
    _device_only("AdvectionRK45")


def AdvectionDiffusionM1(particles, fieldset):  # reference kernels/_advectiondiffusion.py:21-66
    """2-D advection-diffusion, Milstein scheme of first order; needs scalar fields Kh_zonal and Kh_meridional on the
    fieldset's grid (``FieldSet.add_field``) and the context value ``dres`` (finite-difference step for the Kh gradients)."""
    _device_only("AdvectionDiffusionM1")


def AdvectionDiffusionEM(particles, fieldset):  # reference kernels/_advectiondiffusion.py:69-117
    """2-D advection-diffusion, Euler-Maruyama scheme; same requirements as AdvectionDiffusionM1."""
    _device_only("AdvectionDiffusionEM")


def DiffusionUniformKh(particles, fieldset):  # reference kernels/_advectiondiffusion.py:120-153
    """Uniform-Kh Brownian displacement; needs constant fields Kh_zonal and Kh_meridional."""
    _device_only("DiffusionUniformKh")


def DeleteParticle(particles, fieldset):
    """Error handler of the reference's tests (tests/common_kernels.py:12-13): every particle whose
    state is an error (>= 50) is marked Delete.  Must be the LAST kernel of the list."""
    _device_only("DeleteParticle")


# scheme ids of include/parcels_b200.h (enum pb_scheme)
SCHEMES = {"_none": 0, "AdvectionEE": 1, "AdvectionRK2": 2, "AdvectionRK2_3D": 3, "AdvectionRK4": 4, "AdvectionRK4_3D": 5}
SCHEMES_3D = {"AdvectionRK2_3D", "AdvectionRK4_3D"}
ADVDIFF = {"AdvectionDiffusionM1": 0, "AdvectionDiffusionEM": 1}  # enum pb_advdiff_scheme: own entry point (pb_advect_diffusion)
RK45 = 6  # PB_ADVECTION_RK45: own entry point (pb_advect_rk45), per-particle dt / next_dt
