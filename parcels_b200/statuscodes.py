"""Particle status codes and the exceptions mapped to them.

Same integers and class names as the reference (``_core/statuscodes.py:19-128``): the device
writes these ints into ``pset._data["state"]``, and ``ParticleSet.execute`` raises the mapped
exception exactly where ``Kernel.execute`` does (``_core/kernel.py:31-38,239-245``).
"""

from __future__ import annotations

__all__ = [
    "AllParcelsErrorCodes",
    "FieldInterpolationError",
    "FieldOutOfBoundError",
    "FieldOutOfBoundSurfaceError",
    "GeneralError",
    "GridSearchingError",
    "OutsideTimeInterval",
    "StatusCode",
]


class StatusCode:
    Success = 0
    EndofLoop = 1
    Evaluate = 10
    Repeat = 20
    Delete = 30
    StopExecution = 40
    StopAllExecution = 41
    Error = 50
    ErrorInterpolation = 51
    ErrorGridSearching = 52
    ErrorOutOfBounds = 60
    ErrorThroughSurface = 61
    ErrorOutsideTimeInterval = 70


class FieldInterpolationError(RuntimeError):
    pass


class FieldOutOfBoundError(RuntimeError):
    pass


class FieldOutOfBoundSurfaceError(RuntimeError):
    pass


class GridSearchingError(RuntimeError):
    pass


class GeneralError(RuntimeError):
    pass


class OutsideTimeInterval(RuntimeError):
    pass


AllParcelsErrorCodes = {
    FieldInterpolationError: StatusCode.ErrorInterpolation,
    FieldOutOfBoundError: StatusCode.ErrorOutOfBounds,
    FieldOutOfBoundSurfaceError: StatusCode.ErrorThroughSurface,
    GridSearchingError: StatusCode.ErrorGridSearching,
    OutsideTimeInterval: StatusCode.ErrorOutsideTimeInterval,
    GeneralError: StatusCode.Error,
}


def raise_for_state(code, z, y, x, t):
    """Message formats of ``_core/statuscodes.py:44-104``; order of checks is the caller's."""
    if code == StatusCode.ErrorOutsideTimeInterval:
        raise OutsideTimeInterval(f"Field sampled outside time domain at time {t}.")
    if code == StatusCode.ErrorOutOfBounds:
        raise FieldOutOfBoundError(f"Field sampled out-of-bound, at (z={z}, y={y}, x={x})")
    if code == StatusCode.ErrorThroughSurface:
        raise FieldOutOfBoundSurfaceError(f"Field sampled out-of-bound at the surface, at (z={z}, y={y}, x={x})")
    if code == StatusCode.ErrorInterpolation:
        raise FieldInterpolationError(f"Field interpolation returned NaN at (z={z}, y={y}, x={x})")
    if code == StatusCode.ErrorGridSearching:
        raise GridSearchingError(f"Grid searching failed at (z={z}, y={y}, x={x})")
    raise GeneralError(f"General error occurred at (z={z}, y={y}, x={x})")


# order of `ErrorsToThrow` in _core/kernel.py:31-38
ERRORS_TO_THROW = (
    StatusCode.ErrorOutsideTimeInterval,
    StatusCode.ErrorOutOfBounds,
    StatusCode.ErrorThroughSurface,
    StatusCode.ErrorInterpolation,
    StatusCode.ErrorGridSearching,
    StatusCode.Error,
)


class ParticleSetWarning(UserWarning):
    """Issues in the construction of a ParticleSet (reference _core/warnings.py:14-17)."""


class KernelWarning(RuntimeWarning):
    """Warning that a kernel set a default for a missing setting (reference _core/warnings.py)."""
