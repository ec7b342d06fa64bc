"""CPU test (build container only): the adapter reads a live reference FieldSet correctly."""

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle_run import load_case

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not present")


def test_from_parcels_reads_reference_fieldset():
    from parcels_b200 import adapter

    c = load_case("diffusion")
    ref = rh.build_fieldset(lon=c["lon"], lat=c["lat"], depth=c["depth"], times=c["times"], U=c["U"], V=c["V"], W=c["W"],
                            mesh=c["mesh"], constants=c["constants"])  # fmt: skip
    fs = adapter.from_parcels(ref)
    np.testing.assert_array_equal(fs.grid.lon, c["lon"])
    np.testing.assert_array_equal(fs.grid.depth, c["depth"])
    assert fs.grid.is_spherical() and fs.grid.deg2m == ref.U.grid.deg2m
    assert (fs.grid.xdim, fs.grid.ydim, fs.grid.zdim) == (ref.U.grid.xdim, ref.U.grid.ydim, ref.U.grid.zdim)
    np.testing.assert_array_equal(fs._time_s, c["times"])
    np.testing.assert_array_equal(fs.W.data, c["W"])
    assert fs.constants == {"Kh_zonal": 100.0, "Kh_meridional": 50.0}
    assert len(fs.gridset) == len(ref.gridset) == 2
    rp = rh.make_pset(ref, x=c["x"], y=c["y"], z=c["z"], t=c["t"])
    ps = adapter.pset_from_parcels(rp, fs)
    assert ps._data is rp._data and ps._data["ei"].shape == (len(c["x"]), 2)
