"""Device side of the reference's tests/test_spatialhash.py:125-181 (a NaN node invalidates the four faces around it): the table
built ON THE DEVICE from the per-face boxes (csrc/hashbuild.cu) equals the host restatement entry for entry, queries at the centres
of the four faces give GRID_SEARCH_ERROR, every other cell centre resolves to its own cell without a hint -- flat and spherical.
Runs on a GPU, or on the host simulation (PB_LIB=oracle/_build/hostsim/libparcels_b200_hostsim.so PB_HOSTSIM_TEST=1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import analytic as A
import parcels_b200 as pb
from parcels_b200.spatialhash import build_spatial_hash

ok = True
for mesh in ("flat", "spherical"):
    lon, lat = A.rotated_grid()
    clat, clon, jj, ii = A.cell_centers(lon, lat)
    lon, lat = lon.copy(), lat.copy()
    nj, ni = 10, 10
    lon[nj, ni] = lat[nj, ni] = np.nan
    nfx = lon.shape[1] - 1
    invalid = [j * nfx + i for j, i in ((nj - 1, ni - 1), (nj - 1, ni), (nj, ni - 1), (nj, ni))]
    z = np.zeros((1, 1) + lon.shape, dtype=np.float32)
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, U=z, V=z, mesh=mesh, interp_method="cgrid_velocity")
    eng = fs.engine(0)
    dev, host = eng.hash_table(), build_spatial_hash(lon, lat, mesh == "spherical", table=True)
    same = all(np.array_equal(dev[k], host[k]) for k in ("keys", "starts", "counts", "faces"))
    absent = not np.isin(invalid, dev["faces"]).any()
    bad = np.isin(jj * nfx + ii, invalid)
    res = []
    for f32 in (False, True):
        u, v, w, ei, st = eng.sample_velocity(0.0, 0.0, clat, clon, three_d=False, positions_are_f32=f32, no_hint=True)
        res.append((st[bad] == pb.StatusCode.ErrorGridSearching).all() and (st[~bad] == pb.StatusCode.Evaluate).all()
                   and np.array_equal(ei[~bad], (jj * nfx + ii)[~bad]))  # fmt: skip
    print(f"{mesh}: device table == host table {same}, NaN faces absent {absent}, queries f64 / f32 {res}")
    ok &= same and absent and all(res)
    fs.release()
print("PASS nan node" if ok else "FAIL")
sys.exit(0 if ok else 1)
