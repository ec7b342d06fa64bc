"""GPU parity of ONE VectorField.eval (pb_sample_velocity) against the oracle's eval_uvw:
cell indices and error states bit-exact, velocities to a stated relative tolerance."""

import numpy as np
import pytest

import cases
from engine_run import make_fieldset
from oracle import parcels_oracle as po
from oracle_run import load_case, oracle_fieldset

pytestmark = pytest.mark.gpu

NAMES = ["c2_small", "flat_f32c_f64d", "all_f32", "c1_peninsula", "cgrid_rect_3d", "cgrid_rect_sph", "curv_flat_2d",
         "curv_sph_2d", "curv_sph_3d", "curv_sph_f32", "freeslip_3d", "partialslip_sph", "nearest_3d", "freeslip_surface"]  # fmt: skip


def _oracle_eval(ofs, t, z, y, x, three_d, hint):
    n = len(x)
    pd = po.create_particle_data(np.zeros(n), np.zeros(n), np.zeros(n), 0.0, ngrids=ofs.ngrids)
    pd["ei"][:, -1] = hint
    view = po.View(pd, np.ones(n, dtype=bool))
    out = po.eval_uvw(ofs, t, z, y, x, view, three_d)
    return out, pd["ei"][:, -1].copy(), pd["state"].copy()


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("f32", [True, False])
def test_single_eval_matches_oracle(name, f32):
    c = load_case(name)
    fs = make_fieldset(c)
    ofs = oracle_fieldset(c)
    three_d = c["W"] is not None
    dt = np.float32 if f32 else np.float64
    x, y, z = (np.asarray(c[k]).astype(dt) for k in "xyz")
    tmax = 0.0 if c["times"] is None else float(c["times"][-1])
    t = np.full(len(x), 0.37 * tmax)
    eng = fs.engine(0)
    # first eval of a fresh ParticleSet: ei = 0 for everyone => no hint (curvilinear), then with the found cells as hint
    (ou, ov, *ow), oei, ost = _oracle_eval(ofs, t, z, y, x, three_d, np.zeros(len(x), dtype=np.int32))
    u, v, w, ei, st = eng.sample_velocity(t, z, y, x, three_d=three_d, positions_are_f32=f32, ei_hint=np.zeros(len(x), np.int32),
                                          no_hint=True)  # fmt: skip
    np.testing.assert_array_equal(ei, oei)
    np.testing.assert_array_equal(st, ost)
    scale = max(np.abs(ou).max(), np.abs(ov).max())
    np.testing.assert_allclose(u, ou, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(v, ov, rtol=1e-5, atol=1e-6 * scale)
    # second eval, slightly displaced, hinted with the cells just found
    x2 = (x + dt(0.01) * (x.max() - x.min()) / 30).astype(dt)
    (ou, ov, *ow), oei2, ost2 = _oracle_eval(ofs, t, z, y, x2, three_d, oei)
    u, v, w, ei2, st2 = eng.sample_velocity(t, z, y, x2, three_d=three_d, positions_are_f32=f32, ei_hint=oei)
    np.testing.assert_array_equal(ei2, oei2)
    np.testing.assert_array_equal(st2, ost2)
    np.testing.assert_allclose(u, ou, rtol=1e-5, atol=1e-6 * scale)
    np.testing.assert_allclose(v, ov, rtol=1e-5, atol=1e-6 * scale)
    if three_d:
        np.testing.assert_allclose(w, ow[0], rtol=1e-5, atol=1e-6 * np.abs(ow[0]).max())


@pytest.mark.parametrize("name", ["c2_small", "flat_f32c_f64d", "curv_sph_2d", "diffusion"])
def test_populate_indices_matches_oracle_search(name):
    """ParticleSet.populate_indices (reference _core/particleset.py:252-262): ei[:, i] = ravel_index(grid_i.search(z, y, x))
    for every grid of the gridset, the search of the fieldset's grid done on the device."""
    import parcels_b200 as pb

    c = load_case(name)
    fs = make_fieldset(c)
    ofs = oracle_fieldset(c)
    ps = pb.ParticleSet(fs, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    ps.populate_indices()
    d = ps._data
    (zi, _), (yi, _), (xi, _) = po.grid_search(ofs.grid, d["z"], d["y"], d["x"], None)
    np.testing.assert_array_equal(d["ei"][:, 0], po.ravel_index(ofs.grid, zi, yi, xi).astype(np.int32))
    assert d["ei"].shape[1] == ofs.ngrids and not d["ei"][:, 1:].any()
