"""Analytic flows with closed-form trajectories, restated from the formulas of the reference's datasets
(_datasets/structured/generated.py:94-202) and tests (tests/test_advection.py:254-351)."""

import numpy as np

F, U0, UG = 1.0e-4, 0.3, 0.04
GAMMA, GAMMA_G = 1.0 / (2.89 * 86400), 1.0 / (28.9 * 86400)


def moving_eddy(xdim=2, ydim=2):
    """spatially uniform flow rotating in time: U = u_g + (u_0-u_g) cos(f t), V = -(u_0-u_g) sin(f t); 1-minute levels for 7 h"""
    lon = np.linspace(0, 25000, xdim, dtype=np.float32)
    lat = np.linspace(0, 25000, ydim, dtype=np.float32)
    t = np.arange(0, 7 * 3600, 60, dtype=np.float64)
    U = np.zeros((t.size, 1, ydim, xdim), dtype=np.float32)
    V = np.zeros_like(U)
    U[:] = (UG + (U0 - UG) * np.cos(F * t))[:, None, None, None]
    V[:] = (-(U0 - UG) * np.sin(F * t))[:, None, None, None]
    return dict(lon=lon, lat=lat, depth=np.array([0.0], dtype=np.float32), time=t, U=U, V=V)


def moving_eddy_truth(x0, y0, t):
    return x0 + UG * t + (U0 - UG) / F * np.sin(F * t), y0 - (U0 - UG) / F * (1 - np.cos(F * t))


def decaying_eddy(xdim=2, ydim=2):
    lon = np.linspace(0, 20000, xdim, dtype=np.float32)
    lat = np.linspace(5000, 12000, ydim, dtype=np.float32)
    t = np.arange(0, 25 * 3600, 120, dtype=np.float64)
    U = np.zeros((t.size, 1, ydim, xdim), dtype=np.float32)
    V = np.zeros_like(U)
    U[:] = (UG * np.exp(-GAMMA_G * t) + (U0 - UG) * np.exp(-GAMMA * t) * np.cos(F * t))[:, None, None, None]
    V[:] = (-(U0 - UG) * np.exp(-GAMMA * t) * np.sin(F * t))[:, None, None, None]
    return dict(lon=lon, lat=lat, depth=np.array([0.0], dtype=np.float32), time=t, U=U, V=V)


def decaying_eddy_truth(x0, y0, t):
    lon = (x0 + (UG / GAMMA_G) * (1 - np.exp(-GAMMA_G * t))
           + F * ((U0 - UG) / (F**2 + GAMMA**2)) * ((GAMMA / F) + np.exp(-GAMMA * t) * (np.sin(F * t) - (GAMMA / F) * np.cos(F * t))))  # fmt: skip
    lat = y0 - ((U0 - UG) / (F**2 + GAMMA**2)) * F * (1 - np.exp(-GAMMA * t) * (np.cos(F * t) + (GAMMA / F) * np.sin(F * t)))
    return lon, lat


def rotated_grid(ny=60, nx=30):
    """the reference's `2d_left_rotated` fixture mesh (_datasets/structured/generic.py:13-22): integer nodes rotated by -pi/24"""
    LON, LAT = np.meshgrid(np.arange(nx), np.arange(ny))
    a = -np.pi / 24
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    LON, LAT = np.einsum("ji, mni -> jmn", R, np.dstack([LON, LAT]))
    return LON, LAT


def cell_centers(lon, lat):
    clon = 0.25 * (lon[:-1, :-1] + lon[:-1, 1:] + lon[1:, 1:] + lon[1:, :-1])
    clat = 0.25 * (lat[:-1, :-1] + lat[:-1, 1:] + lat[1:, 1:] + lat[1:, :-1])
    jj, ii = np.meshgrid(np.arange(lon.shape[0] - 1), np.arange(lon.shape[1] - 1), indexing="ij")
    return clat.ravel(), clon.ravel(), jj.ravel(), ii.ravel()


def stommel_gyre(xdim=200, ydim=200):
    """Western-boundary-current gyre on an A-grid (reference _datasets/structured/generated.py:301-357): sea-surface height
    P = (1 - exp(-x/eps) - x) pi sin(pi y) s and its geostrophic velocities, float32 arrays on a 10 000 km square."""
    a = b = 10000 * 1e3
    s = 0.05
    lon = np.linspace(0, a, xdim, dtype=np.float32)
    lat = np.linspace(0, b, ydim, dtype=np.float32)
    eps = (1 / (11.6 * 86400)) / (2e-11 * a)
    xi, yi = (lon / a).astype(np.float64)[None, :], (lat / b).astype(np.float64)[:, None]
    g = 1 - np.exp(-xi / eps) - xi
    P = (g * np.pi * np.sin(np.pi * yi) * s).astype(np.float32)
    U = (-g * np.pi**2 * np.cos(np.pi * yi) * s).astype(np.float32)
    V = ((np.exp(-xi / eps) / eps - 1) * np.pi * np.sin(np.pi * yi) * s).astype(np.float32)
    return dict(lon=lon, lat=lat, U=U[None, None], V=V[None, None], P=P[None, None])
