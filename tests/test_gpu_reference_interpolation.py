"""The reference's own ``tests/test_interpolation.py`` known answers, transcribed (same names, same expected values), evaluated
through this package's public sampling API (`fieldset.P[t, z, y, x]`, `fieldset.UV[t, z, y, x]`) on the device.  The v3 regression
test of that file is tests/test_gpu_parity.py::test_v3_*; its corner-gather tests are about an internal helper of the reference."""

import numpy as np
import pytest

import parcels_b200 as pb

pytestmark = pytest.mark.gpu


def _data():
    z0 = np.array([[0.0, 1.0, 2.0, 3.0], [2.0, 3.0, 4.0, 5.0], [4.0, 5.0, 6.0, 7.0], [6.0, 7.0, 8.0, 9.0]])  # x: +1, y: +2
    spatial = np.array([z0, z0 + 3, z0 + 6, z0 + 9])  # z: +3
    return np.array([spatial, spatial + 10, spatial + 20])  # t: +10


def _fieldset(P, method="linear", mesh="flat", uv=None, interp="linear"):
    n = np.array([0.0, 1.0, 2.0, 3.0])
    U = np.zeros_like(P) if uv is None else uv
    fs = pb.FieldSet.from_arrays(lon=n, lat=n, depth=n, time=np.array([np.timedelta64(t, "s") for t in [0, 2, 4]]), U=U, V=U.copy(), mesh=mesh,
                                 interp_method=interp)  # fmt: skip
    fs.add_field("P", P, interp_method=method)
    return fs


@pytest.mark.parametrize("method, t, z, y, x, expected", [
    pytest.param("linear", [0, 1], [0, 0], [0.49, 0.49], [0.51, 0.51], [1.49, 6.49], id="Linear-1"),
    pytest.param("linear", 1, 2.5, 0.49, 0.51, 13.99, id="Linear-2"),
    pytest.param("linear", [0, 1, 1], [0, 0, 2.5], [0.49, 0.49, 0.49], [0.51, 0.51, 0.51], [1.49, 6.49, 13.99], id="Linear-3"),
    pytest.param("linear_invdist_land", 1, 2.5, 0.49, 0.51, 13.99, id="LinearInvDistLand"),
    pytest.param("nearest", [0, 3], [0.2, 0.2], [0.2, 0.2], [0.51, 0.51], [1.0, 16.0], id="Nearest"),
])  # fmt: skip
def test_raw_2d_interpolation(method, t, z, y, x, expected):
    fs = _fieldset(_data(), method)
    value = fs.P[np.asarray(t, dtype=np.float64), np.asarray(z, dtype=np.float64), np.asarray(y), np.asarray(x)]
    np.testing.assert_equal(value, np.atleast_1d(expected))


@pytest.mark.parametrize("mesh", ["flat", "spherical"])
@pytest.mark.parametrize("func, t, z, y, x, expected", [
    ("partialslip", 1, 0, 0, 0.0, [[1.0], [1.0]]),
    ("freeslip", 1, 0, 0.5, 1.5, [[1.0], [0.5]]),
    ("partialslip", 1, 0, 2.5, 1.5, [[0.75], [0.5]]),
    ("freeslip", 1, 0, 2.5, 1.5, [[1.0], [0.5]]),
    ("partialslip", 1, 0, 1.5, 0.5, [[0.5], [0.75]]),
    ("freeslip", 1, 0, 1.5, 0.5, [[0.5], [1.0]]),
    ("freeslip", [1, 0], [0, 2], [1.5, 1.5], [2.5, 0.5], [[0.5, 0.5], [1.0, 1.0]]),
])  # fmt: skip
def test_spatial_slip_interpolation(func, t, z, y, x, expected, mesh):
    data = np.ones_like(_data())
    data[:, :, 1:3, 1:3] = 0.0  # zero land value to test spatial slip
    fs = _fieldset(data, mesh=mesh, uv=data, interp=func)  # U = V = the field, UV with the slip interpolator
    velocities = fs.UV[np.asarray(t, dtype=np.float64), np.asarray(z, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64)]
    expected = np.array(expected)
    if mesh == "spherical":
        expected[0] = expected[0] / (1852 * 60.0 * np.cos(np.radians(y)))
        expected[1] = expected[1] / (1852 * 60.0)
    np.testing.assert_array_almost_equal(np.array(velocities), expected)


@pytest.mark.parametrize("t, z, y, x, expected", [
    (1, 0, 0.5, 0.5, 1.0),
    (1, 0, 1.5, 1.5, 0.0),
    ([0, 1], [0, 2], [0.5, 0.5], [0.5, 0.5], 1.0),
    ([0, 1], [0, 2], [0.5, 1.5], [0.5, 1.5], [1.0, 0.0]),
])  # fmt: skip
def test_invdistland_interpolation(t, z, y, x, expected):
    data = np.ones_like(_data())
    data[:, :, 1:3, 1:3] = 0
    fs = _fieldset(data, "linear_invdist_land")
    value = fs.P[np.asarray(t, dtype=np.float64), np.asarray(z, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64)]
    np.testing.assert_array_almost_equal(value, np.broadcast_to(expected, np.shape(value)))


@pytest.mark.parametrize("mesh", ["spherical", "flat"])
def test_interpolation_mesh_type(mesh):
    # simple_UV_dataset(mesh=mesh) with U = 1
    max_lon, max_lat = (180.0, 90.0) if mesh == "spherical" else (1e6, 1e6)
    dims = (360, 2, 30, 4)
    time = np.datetime64("2000-01-01") + (np.arange(dims[0]) * (366 * 86400 / (dims[0] - 1))).astype("timedelta64[s]")
    fieldset = pb.FieldSet.from_arrays(lon=np.linspace(-max_lon, max_lon, 4), lat=np.linspace(-max_lat, max_lat, 30), depth=np.linspace(0, 1, 2),
                                       time=time, U=np.ones(dims), V=np.zeros(dims), mesh=mesh)  # fmt: skip
    lat = 30.0
    time = 0.0
    u_expected = 1.0 if mesh == "flat" else 1.0 / (1852 * 60 * np.cos(np.radians(lat)))
    with pytest.warns(RuntimeWarning, match="Sampling of velocities should normally be done"):
        assert fieldset.U.eval(time, 0, lat, 0) == 1.0
        assert fieldset.V[time, 0, lat, 0] == 0.0
    u, v = fieldset.UV[time, 0, lat, 0]
    assert np.isclose(u, u_expected, atol=1e-7)
    assert v == 0.0


def _ramp_fieldset():
    """lon = [1, 2, 3, 4, 5] (the array of the reference's tests/test_xgrid.py::test_search_1d_array*), P = lon: the sampled value of
    XLinear is lon[xi] + xsi * (lon[xi + 1] - lon[xi]), the returned cell index is xi (one row of cells)."""
    lon = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    P = np.broadcast_to(lon, (1, 1, 2, 5)).copy()
    fs = pb.FieldSet.from_arrays(lon=lon, lat=np.array([0.0, 1.0]), U=np.zeros_like(P), V=np.zeros_like(P), mesh="flat")
    fs.add_field("P", P)
    return fs, lon


@pytest.mark.parametrize("x, expected_xi, expected_xsi", [((1.1, 2.1), (0, 1), (0.1, 0.1)), (2.1, 1, 0.1), (3.1, 2, 0.1), (4.5, 3, 0.5)])
def test_search_1d_array(x, expected_xi, expected_xsi):
    fs, lon = _ramp_fieldset()
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    val, ei, st = fs.engine(0).sample_scalar(fs.P._slot, "linear", np.zeros_like(x), np.zeros_like(x), np.full_like(x, 0.5), x, positions_are_f32=False, ei_hint=None)
    np.testing.assert_array_equal(ei, np.atleast_1d(expected_xi))
    np.testing.assert_allclose(val - lon[ei], np.atleast_1d(expected_xsi))
    assert np.all(st == pb.StatusCode.Evaluate)


@pytest.mark.parametrize("x, expected_state", [(-0.1, pb.StatusCode.Evaluate), (6.5, pb.StatusCode.ErrorOutOfBounds), ((-0.1, 2.5), None), ((6.5, 1), None)])
def test_search_1d_array_out_of_bounds(x, expected_state):
    """LEFT_OUT_OF_BOUNDS (-2) is not an error state in the reference for X / Y (field.py:327-356), RIGHT_OUT_OF_BOUNDS (-1) is
    ErrorOutOfBounds; either way the sampled value is 0 (field.py:359-370) and in-bounds samples of the same call are unaffected."""
    fs, lon = _ramp_fieldset()
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    val, ei, st = fs.engine(0).sample_scalar(fs.P._slot, "linear", np.zeros_like(x), np.zeros_like(x), np.full_like(x, 0.5), x, positions_are_f32=False, ei_hint=None)
    inside = (x >= 1) & (x <= 5)
    assert np.all(val[~inside] == 0.0)
    np.testing.assert_allclose(val[inside], x[inside])
    np.testing.assert_array_equal(st[x > 5], pb.StatusCode.ErrorOutOfBounds)
    np.testing.assert_array_equal(st[x < 1], pb.StatusCode.Evaluate)
    if expected_state is not None:
        assert st[0] == expected_state


def test_grid_indexing_fpoints():
    """reference tests/test_index_search.py::test_grid_indexing_fpoints on the device: on the `2d_left_unrolled_cone` curvilinear
    mesh (reference _datasets/structured/generic.py:76-100, flat), every F-point nudged by 1e-5 is found in the cell it is the
    lower-left corner of (or its left / lower neighbour when the nudge barely crosses), and that cell's corners bracket it."""
    X, Y = 30, 60
    XG, YG = np.arange(X), np.arange(Y) * 0.25
    LON, LAT = np.meshgrid(XG, YG)
    pivot = (-10.0, 0.0)
    r = np.sqrt((LON - pivot[0]) ** 2 + (LAT - pivot[1]) ** 2) * 1.2
    theta = np.arctan2(LAT - pivot[1], XG.min() - pivot[0]) * 1.2
    lon, lat = r * np.cos(theta) + pivot[0], r * np.sin(theta) + pivot[1]
    z = np.zeros((1, 1, Y, X))
    fs = pb.FieldSet.from_arrays(lon=lon, lat=lat, U=z, V=z, mesh="flat", interp_method="cgrid_velocity")
    ydim, xdim = Y - 1, X - 1  # cell counts
    jj, ii = np.meshgrid(np.arange(ydim - 1), np.arange(xdim - 1), indexing="ij")
    jj, ii = jj.ravel(), ii.ravel()
    x, y = lon[jj, ii] + 0.00001, lat[jj, ii] + 0.00001
    u, v, w, ei, st = fs.engine(0).sample_velocity(0.0, 0.0, y, x, three_d=False, positions_are_f32=False, no_hint=True)
    assert np.all(st == pb.StatusCode.Evaluate)
    yi, xi = ei // xdim, ei % xdim
    # the reference accepts the neighbour below / to the left when eta / xsi > 0.9 there: the cell found must be one of those
    assert np.all((yi == jj) | (yi == jj - 1)) and np.all((xi == ii) | (xi == ii - 1))
    cell_lon = np.stack([lon[yi, xi], lon[yi, xi + 1], lon[yi + 1, xi + 1], lon[yi + 1, xi]])
    cell_lat = np.stack([lat[yi, xi], lat[yi, xi + 1], lat[yi + 1, xi + 1], lat[yi + 1, xi]])
    assert np.all((x > cell_lon.min(axis=0)) & (x < cell_lon.max(axis=0)))
    assert np.all((y > cell_lat.min(axis=0)) & (y < cell_lat.max(axis=0)))
