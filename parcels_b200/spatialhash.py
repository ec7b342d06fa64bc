"""One-time construction of the spatial-hash table the device query walks.

Same table as the reference's ``SpatialHash`` (``_core/spatialhash.py:45-387``): per-face bounding
boxes (unit-sphere xyz on spherical meshes, lon/lat on flat ones) quantised to <= 10 bits per axis,
one (Morton key, face) entry per hash cell a box overlaps, sorted by (key, face) and stored CSR --
so the device sees candidates in the reference's order ("first containing face wins").  This is
grid *setup* (like uploading the coordinates); the query runs in ``csrc/cgrid.cu``.

Split: the float part (xyz in the coordinate dtype, per-face min/max, quantisation, bitwidth budget
search -- O(faces), and NumPy's sin/cos are what fix the reference's boxes) runs here;
``build_spatial_hash(..., table=False)`` stops after it and the O(entries) integer part (expansion,
Morton encode, (key, face) sort, CSR) runs on the device (``csrc/hashbuild.cu``).  ``table=True`` also
does the integer part in NumPy -- the host restatement the GPU tests compare the device table with.
"""

from __future__ import annotations

import numpy as np

ENTRIES_PER_FACE = 16  # _core/spatialhash.py:24-26
ENTRY_BUDGET_MIN = 2**22
MAX_BITWIDTH = 1023


def _xyz(lat_deg, lon_deg):
    lat, lon = np.deg2rad(lat_deg), np.deg2rad(lon_deg)  # keeps the coordinate dtype, like the reference
    return np.cos(lon) * np.cos(lat), np.sin(lon) * np.cos(lat), np.sin(lat)


def _quantize(v, lo, hi, bitwidth):
    d = hi - lo
    with np.errstate(invalid="ignore"):
        vn = np.where(d != 0, (np.asarray(v) - lo) / d, 0.0)
        return np.clip(vn * bitwidth, 0, bitwidth).astype(np.uint32)


def _dilate(n):
    n = np.asarray(n, dtype=np.uint32) & np.uint32(0x3FF)
    for shift, mask in ((16, 0xFF0000FF), (8, 0x0300F00F), (4, 0x030C30C3), (2, 0x09249249)):
        n = (n | (n << np.uint32(shift))) & np.uint32(mask)
    return n


def _face_minmax(a):
    c = np.stack((a[:-1, :-1], a[:-1, 1:], a[1:, 1:], a[1:, :-1]), axis=-1)
    return c.min(axis=-1), c.max(axis=-1)


def build_spatial_hash(lon2d: np.ndarray, lat2d: np.ndarray, spherical: bool, table: bool = True) -> dict:
    """Returns dict(box f64[6], bitwidth int, qbox u64[faces], n_entries int) and, with ``table=True``, the CSR table
    itself: keys u32, starts i64, counts i64, faces u32."""
    if spherical:
        x, y, z = _xyz(lat2d, lon2d)
        box = (np.nanmin(x), np.nanmax(x), np.nanmin(y), np.nanmax(y), np.nanmin(z), np.nanmax(z))
    else:
        x, y, z = lon2d, lat2d, None
        box = (np.nanmin(x), np.nanmax(x), np.nanmin(y), np.nanmax(y), 0.0, 0.0)
    xl, xh = _face_minmax(x)
    yl, yh = _face_minmax(y)
    zl, zh = _face_minmax(z) if z is not None else (np.zeros_like(xl), np.zeros_like(xl))
    valid = ~(np.isnan(xl) | np.isnan(xh) | np.isnan(yl) | np.isnan(yh) | np.isnan(zl) | np.isnan(zh)).ravel()

    def boxes(bw):
        lo = [_quantize(a, box[2 * k], box[2 * k + 1], bw).ravel().astype(np.int64) for k, a in enumerate((xl, yl, zl))]
        hi = [_quantize(a, box[2 * k], box[2 * k + 1], bw).ravel().astype(np.int64) for k, a in enumerate((xh, yh, zh))]
        return lo, hi

    def total(bw):
        lo, hi = boxes(bw)
        n = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
        return int(np.where(valid, n, 0).sum())

    bitwidth = MAX_BITWIDTH
    budget = max(ENTRIES_PER_FACE * xl.size, ENTRY_BUDGET_MIN)
    if total(bitwidth) > budget:  # largest bitwidth within budget (binary search, spatialhash.py:212-228)
        lo_b, hi_b = 1, bitwidth
        while lo_b < hi_b:
            mid = (lo_b + hi_b + 1) // 2
            if total(mid) <= budget:
                lo_b = mid
            else:
                hi_b = mid - 1
        bitwidth = lo_b
    lo, hi = boxes(bitwidth)
    nx, ny, nz = (hi[k] - lo[k] + 1 for k in range(3))
    per_face = np.where(valid, nx * ny * nz, 0)
    nent = int(per_face.sum())
    # per-face quantised bounding box, packed 6 x 10 bits: a face is listed under hash cell (qx, qy, qz) iff the
    # cell lies inside this box -- the device builds the table from it and decides table membership without walking it
    qbox = (lo[0] | (hi[0] << 10) | (lo[1] << 20) | (hi[1] << 30) | (lo[2] << 40) | (hi[2] << 50)).astype(np.uint64)
    qbox[~valid] = np.uint64(1023)  # lo = 1023 > hi = 0: empty
    head = dict(box=np.array([float(b) for b in box], dtype=np.float64), bitwidth=int(bitwidth), qbox=np.ascontiguousarray(qbox),
                n_entries=nent)  # fmt: skip
    if not table:
        return head
    face = np.repeat(np.arange(xl.size, dtype=np.uint32), per_face)
    intra = np.arange(nent, dtype=np.int64) - np.repeat(np.concatenate(([0], np.cumsum(per_face)))[:-1], per_face)
    nynz, nzr = np.repeat(ny * nz, per_face), np.repeat(nz, per_face)
    cx = np.repeat(lo[0], per_face) + intra // nynz
    cy = np.repeat(lo[1], per_face) + (intra % nynz) // nzr
    cz = np.repeat(lo[2], per_face) + (intra % nynz) % nzr
    code = (_dilate(cz) << np.uint32(2)) | (_dilate(cy) << np.uint32(1)) | _dilate(cx)
    packed = (code.astype(np.uint64) << np.uint64(32)) | face
    packed.sort()
    faces = packed.astype(np.uint32)
    codes = (packed >> np.uint64(32)).astype(np.uint32)
    starts = np.concatenate(([0], np.flatnonzero(codes[1:] != codes[:-1]) + 1)).astype(np.int64)
    counts = np.diff(np.concatenate((starts, [codes.size]))).astype(np.int64)
    return dict(head, keys=np.ascontiguousarray(codes[starts]), starts=starts, counts=counts, faces=np.ascontiguousarray(faces))
