"""CPU-side checks of what surrounds the hot path: the dew-point oracle pinned to the reference-generated golden vectors
(and to the reference C compiled in place where it is present), the NetCDF-3 float32 sink, the date-indexed queues."""
import os

import numpy as np
import pytest


def test_dew_point_oracle_matches_reference_golden(golden):
    from oracle import oracle as orc
    g = golden("dewpt")
    for tol in (0.01, 0.001):
        np.testing.assert_array_equal(orc.dew_point(g["vp"], tol), g["dpt_tol%g" % tol])      # bit for bit
    ref_so = os.path.join(os.path.dirname(orc.__file__), "_ref", "libenvphys_ref.so")
    if os.path.exists(ref_so):
        vp = np.random.default_rng(3).uniform(20.0, 5000.0, 20000)
        np.testing.assert_array_equal(orc.dew_point(vp, 0.01), orc.dew_point(vp, 0.01, impl="ref"))
    assert np.isnan(orc.dew_point(np.array([np.nan, 800.0]), 0.01)[0])
    d = orc.dew_point(np.array([611.0]), 1e-6)
    assert abs(d[0]) < 0.2                                    # ~0 degC at the triple-point pressure


def test_wet_bulb_oracle_matches_reference_golden(golden):
    """oracle/envphys_oracle.c (restatement of iwbt.c:14-134) against vectors produced by the reference's iwbt.c itself."""
    from oracle import oracle as orc
    g = golden("wetbulb")
    for tol in (0.01, 0.05):
        np.testing.assert_array_equal(orc.wet_bulb(g["ta"], g["td"], g["z"], tol), g["tw_tol%g" % tol])      # bit for bit
    ref_so = os.path.join(os.path.dirname(orc.__file__), "_ref", "libenvphys_ref.so")
    if os.path.exists(ref_so):
        rng = np.random.default_rng(8)
        ta = rng.uniform(-30.0, 30.0, 20000)
        td = ta - rng.gamma(1.5, 4.0, 20000)
        z = rng.uniform(0.0, 4000.0, 20000)
        np.testing.assert_array_equal(orc.wet_bulb(ta, td, z, 0.01), orc.wet_bulb(ta, td, z, 0.01, impl="ref"))
    tw = orc.wet_bulb(np.array([10.0, -5.0]), np.array([10.0, -9.0]), np.array([0.0, 2000.0]), 0.001)
    assert abs(tw[0] - 10.0) < 1e-6 and -9.0 < tw[1] < -5.0         # saturated air: wet bulb = air temperature; else in between
    with pytest.raises(RuntimeError):
        orc.wet_bulb(np.array([-300.0]), np.array([-300.0]), np.array([0.0]), 0.01)     # below 0 K: the reference exits


def test_snow_phase_oracle_matches_reference_golden(golden):
    """oracle.snow_phase_and_density (restatement of envphys/snow.py + nasde_model/*.py) against vectors produced by executing
    the reference's own Snow.phase_and_density (tests/golden/make_golden_snow.py): bit for bit, all three models, on the hourly
    precipitation, on a storm total and on an all-dry hour."""
    from oracle import oracle as orc
    g = golden("snow")
    for model in ("susong1999", "piecewise_susong1999", "marks2017"):
        for tag in ("pp", "storm", "zero"):
            t0 = g["t"].copy()
            rho, pcs = orc.snow_phase_and_density(t0, g[tag], model)
            np.testing.assert_array_equal(rho, g["%s_%s_rho" % (model, tag)])
            np.testing.assert_array_equal(pcs, g["%s_%s_pcs" % (model, tag)])
            np.testing.assert_array_equal(t0, g["t"])                  # the caller's temperature is not clipped
    rho, pcs = orc.snow_phase_and_density(np.array([-7.0, -0.25, 0.25, 3.0]), np.array([1.0, 1.0, 1.0, 1.0]), "susong1999")
    np.testing.assert_array_equal(rho, [75.0, 200.0, 250.0, 0.0])
    np.testing.assert_array_equal(pcs, [1.0, 0.75, 0.25, 0.0])
    with pytest.raises(ValueError):
        orc.snow_phase_and_density(np.zeros(2), np.zeros(2), "nope")


def test_wind_direction_and_weights_oracle():
    from oracle import oracle as orc
    az = orc.wind_direction(np.array([0.0, 1.0, 0.0, -1.0, -0.5]), np.array([1.0, 0.0, -1.0, 0.0, 0.5]))
    np.testing.assert_allclose(az, [0.0, 90.0, 180.0, 270.0, 315.0])
    gx, gy = np.meshgrid(np.arange(5.0), np.arange(4.0))
    xy = np.column_stack([gx.ravel(), gy.ravel()])
    uv = np.array([[0.5, 0.5], [3.25, 2.5], [9.0, 9.0]])
    vtx, wts = orc.interp_weights(xy, uv)
    f = 2.0 * xy[:, 0] - 3.0 * xy[:, 1] + 1.0                 # linear field: reproduced exactly inside the hull
    r = orc.grid_interpolate(f, vtx, wts, (1, 3))
    np.testing.assert_allclose(r[0, :2], 2.0 * uv[:2, 0] - 3.0 * uv[:2, 1] + 1.0, rtol=1e-13)
    assert np.isnan(r[0, 2])


def test_netcdf3_sink_roundtrip(tmp_path):
    from scipy.io import netcdf_file
    from smrf_b200.output import BlockSink
    x = 5e5 + 30.0 * np.arange(9)
    y = 4e6 - 30.0 * np.arange(6)
    mask = np.ones((6, 9))
    mask[0, :] = 0.0
    rng = np.random.default_rng(0)
    blk = rng.normal(size=(5, 6, 9))
    s = BlockSink(str(tmp_path), ["air_temp", "precip"], x, y, hours=np.arange(5), mask=mask, mask_output=True,
                  start_date="1998-01-14 15:00:00", units={"air_temp": "degree_Celsius"})
    s("air_temp", (0, 3), (0, 4), blk[0:3, 0:4])               # row slabs and time blocks in any order
    s("air_temp", (0, 3), (4, 6), blk[0:3, 4:6])
    s("air_temp", (3, 5), (0, 6), blk[3:5].astype(np.float32))
    s("precip", (0, 5), (0, 6), blk)
    s.close()
    f = netcdf_file(os.path.join(str(tmp_path), "air_temp.nc"), mmap=False)
    v = f.variables["air_temp"]
    assert v.dimensions == ("time", "y", "x") and v.data.dtype.itemsize == 4 and v.units == b"degree_Celsius"
    np.testing.assert_array_equal(v[:], (blk * mask).astype(np.float32))
    np.testing.assert_array_equal(f.variables["time"][:], np.arange(5, dtype=np.float32))
    assert f.variables["time"].units == b"hours since 1998-01-14 15:00:00"
    np.testing.assert_allclose(f.variables["x"][:], x)


def test_date_queue_protocol():
    import threading
    from smrf_b200.run_loop import DateQueue
    q = DateQueue(timeout=5.0)
    got = []
    th = threading.Thread(target=lambda: got.append(q.get("t1")))
    th.start()
    q.put(["t0", 1])
    q.put(["t1", 2])
    th.join()
    assert got == [2] and q.get("t0") == 1
    with pytest.raises(TimeoutError):
        DateQueue(timeout=0.05).get("never")
