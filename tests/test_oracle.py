"""Pin the CPU oracle against the reference-generated golden vectors (tests/golden/*.npz,
produced by tests/golden/make_golden.py executing /root/reference) and, when available, the
reference kriging C compiled in place (oracle/_ref).  CPU only."""
import warnings

import numpy as np
import pytest

from oracle import build as obuild
from oracle import oracle as orc


@pytest.fixture(scope="module", autouse=True)
def _built():
    obuild.build_oracle()
    obuild.build_ref()


def test_clamp_matches_reference(golden):
    g = golden("clamp")
    np.testing.assert_array_equal(orc.set_min_max(g["a"].copy(), -1.0, 2.0), g["out_both"])
    np.testing.assert_array_equal(orc.set_min_max(g["a"].copy(), 0.0, None), g["out_min"])
    np.testing.assert_array_equal(orc.set_min_max(g["a"].copy(), None, None), g["out_none"])


def test_idw_rme_bitwise(golden):
    g = golden("idw_rme")
    X, Y = np.meshgrid(g["x"], g["y"])
    o = orc.IDW(g["mx"], g["my"], X, Y, mz=g["mz"], GridZ=g["dem"], power=2.0)
    for t, row in enumerate(g["air_temp"]):
        v = orc.set_min_max(o.detrendedIDW(row.copy(), -1), -73.0, 47.0)
        np.testing.assert_array_equal(v, g["out_air_temp"][t])
        np.testing.assert_array_equal(o.pv, g["pv_air_temp"][t])
    for t, row in enumerate(g["vapor_pressure"]):
        v = orc.set_min_max(o.detrendedIDW(row.copy(), -1), 20.0, 5000.0)
        np.testing.assert_array_equal(v, g["out_vapor_pressure"][t])
    for t, row in enumerate(g["cloud_factor"]):
        v = orc.set_min_max(o.calculateIDW(row.copy()), 0.0, 1.0)
        np.testing.assert_array_equal(v, g["out_cloud_factor"][t])


@pytest.mark.parametrize("power", [2.0, 1.5])
def test_idw_synth_bitwise(golden, power):
    g = golden("idw_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    o = orc.IDW(g["mx"], g["my"], X, Y, mz=g["mz"], GridZ=g["dem"], power=power)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for flag in (-1, 0, 1):
            for t, row in enumerate(g["data"]):
                v = o.detrendedIDW(row.copy(), flag)
                np.testing.assert_array_equal(v, g["out_p%g_f%d" % (power, flag)][t])
                np.testing.assert_array_equal(o.pv, g["pv_p%g_f%d" % (power, flag)][t])
        for t, row in enumerate(g["data"]):
            np.testing.assert_array_equal(o.calculateIDW(row.copy()), g["out_p%g_plain" % power][t])


def test_idw_zero_distance_quirk(golden):
    g = golden("idw_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    with np.errstate(all="ignore"):
        o = orc.IDW(g["mx0"], g["my0"], X, Y, mz=g["mz"], GridZ=g["dem"], power=2.0)
        for t, row in enumerate(g["data"][:4]):
            v = o.detrendedIDW(row.copy(), 0)
            np.testing.assert_array_equal(v, g["out_zero_dist"][t])
    # the coincident cell is NaN in the reference whenever that station reports (steps 1, 2)
    assert np.isnan(g["out_zero_dist"][1][7, 5]) and np.isnan(g["out_zero_dist"][2][7, 5])
    assert not np.isnan(g["out_zero_dist"][0][7, 5])      # station 3 is missing at step 0


def test_krige_kat_bitwise(golden):
    g = golden("krige_kat")
    w = np.zeros_like(g["dgrid"])
    act = np.zeros(g["dgrid"].shape, dtype=np.int32)
    orc.call_grid(g["ad"], g["dgrid"], g["sz"], w, 2, active=act)
    np.testing.assert_array_equal(w, g["weights"])
    np.testing.assert_array_equal(act != 0, g["weights"] != 0)
    np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=0, atol=1e-10)
    assert (w >= 0).all()


@pytest.mark.parametrize("nsta,ncell,seed", [(2, 40, 0), (5, 200, 1), (13, 300, 2), (40, 120, 3), (80, 24, 4)])
def test_krige_vs_compiled_reference(nsta, ncell, seed):
    if orc.krige_ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(seed)
    sx, sy = rng.uniform(0, 30000, nsta), rng.uniform(0, 30000, nsta)
    sz = rng.uniform(900, 3000, nsta)
    if nsta > 4:
        sz[1] = -5.0                                   # elevation <= 0 can never be dropped (krige.c:172)
    cx, cy = rng.uniform(-2000, 32000, ncell), rng.uniform(-2000, 32000, ncell)
    ad = orc.station_distances(sx, sy)
    dg = np.sqrt((cx[:, None] - sx[None, :]) ** 2 + (cy[:, None] - sy[None, :]) ** 2)
    w_ref = np.zeros_like(dg)
    w_orc = np.zeros_like(dg)
    orc.call_grid(ad, dg, sz, w_ref, 2, impl="ref")
    orc.call_grid(ad, dg, sz, w_orc, 2, impl="oracle")
    np.testing.assert_array_equal(w_orc, w_ref)


def test_dk_synth_bitwise(golden):
    g = golden("dk_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    o = orc.DK(g["mx"], g["my"], g["mz"], X, Y, g["dem"], {"dk_ncores": 2, "detrend_slope": 1})
    for t, row in enumerate(g["data"]):
        v = orc.set_min_max(o.calculate(row.copy()), 0.0, None)
        np.testing.assert_array_equal(v, g["out"][t])
        np.testing.assert_array_equal(o.nan_val, g["nan_mask"][t])
        np.testing.assert_array_equal(o.pv, g["pv"][t])
        full = np.zeros(X.shape + (len(g["mx"]),))
        full[:, :, ~o.nan_val] = o.weights
        np.testing.assert_array_equal(full, g["weights"][t])


def test_dk_rme_bitwise(golden):
    g = golden("dk_rme")
    X, Y = np.meshgrid(g["x"], g["y"])
    o = orc.DK(g["mx"], g["my"], g["mz"], X, Y, g["dem"], {"dk_ncores": 1, "detrend_slope": 1})
    for t, row in enumerate(g["data"]):
        v = orc.set_min_max(o.calculate(row.copy()), 0.0, None)
        np.testing.assert_array_equal(v, g["out"][t])


def test_grid_synth_bitwise(golden):
    g = golden("grid_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    plan = orc.linear_plan(g["px"], g["py"], X, Y)
    assert (plan["simplex"] < 0).sum() == np.isnan(g["out_plain"][0]).sum() > 0
    for gm in (1, 0):
        cfg = {"grid_local": False, "grid_mask": bool(gm)}
        o = orc.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], mask=g["mask"])
        np.testing.assert_array_equal(o.mask, g["ptmask_%d" % gm])
        for flag in (-1, 0):
            for t, row in enumerate(g["fields"]):
                v = o.detrendedInterpolation(row.copy(), flag, "linear")
                np.testing.assert_array_equal(v, g["out_m%d_f%d" % (gm, flag)][t])
                np.testing.assert_array_equal(o.pv, g["pv_m%d_f%d" % (gm, flag)][t])
                # decomposed restatement (what the CUDA path does) == griddata, bit for bit
                dtrend = row - (g["pz"] * o.pv[0] + o.pv[1])
                v2 = orc.linear_apply(plan, dtrend, X, Y) + o.pv[0] * g["dem"] + o.pv[1]
                np.testing.assert_array_equal(v2, v)
    for t, row in enumerate(g["fields"]):
        np.testing.assert_array_equal(orc.linear_apply(plan, row, X, Y), g["out_plain"][t])


def _local_metadata(g):
    import pandas as pd
    return pd.DataFrame({"utm_x": g["px"], "utm_y": g["py"], "elevation": g["pz"], "latitude": g["lat"],
                         "longitude": g["lon"]})


def test_grid_local_bitwise(golden):
    """detrendedInterpolationLocal (grid.py:115-177, executed by the reference: tests/golden/make_golden_local.py):
    neighbour table, per-point fits with the slope constraint, three interpolations and their combination."""
    g = golden("grid_local")
    X, Y = np.meshgrid(g["x"], g["y"])
    cfg = {"grid_local": True, "grid_local_n": int(g["k"]), "grid_mask": False}
    o = orc.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], metadata=_local_metadata(g))
    np.testing.assert_array_equal(o.nbr, g["nbr"])
    assert np.isnan(g["out_f0"][0]).sum() > 0
    for flag in (-1, 0, 1):
        for t, row in enumerate(g["fields"]):
            np.testing.assert_array_equal(o.detrendedInterpolation(row.copy(), flag, "linear"), g["out_f%d" % flag][t])
    # the constraint bites: flag +1 zeroes the fits with a negative slope (these fields have a lapse rate)
    res, slope, icpt = orc.local_fit(g["fields"][0], g["pz"], g["nbr"], 1)
    res0, slope0, _ = orc.local_fit(g["fields"][0], g["pz"], g["nbr"], 0)
    assert (slope0 < 0).any() and not (slope < 0).any() and np.all(icpt[slope0 < 0] == 0.0)
    assert not np.array_equal(g["out_f1"], g["out_f0"])


def test_grid_cubic_golden(golden):
    """grid_method='cubic' (Clough-Tocher), executed by the reference (tests/golden/make_golden_cubic.py): the oracle's
    GRID calls SciPy exactly where grid.py / utils.py do -> bit-identical; the decomposed restatement the CUDA path
    uses (nine weights per cell, SciPy's own gradient estimator) agrees to rounding."""
    g = golden("grid_cubic")
    X, Y = np.meshgrid(g["x"], g["y"])
    md = _local_metadata(g)
    cfg = {"grid_local": True, "grid_local_n": int(g["k"]), "grid_mask": False}
    o = orc.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], metadata=md)
    for flag in (-1, 0, 1):
        for t, row in enumerate(g["fields"]):
            np.testing.assert_array_equal(o.detrendedInterpolation(row.copy(), flag, "cubic"), g["local_f%d" % flag][t])
    assert not np.array_equal(g["local_f0"], golden("grid_local")["out_f0"])          # cubic differs from linear
    o2 = orc.GRID({"grid_local": False, "grid_mask": True}, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], mask=g["mask"])
    np.testing.assert_array_equal(o2.mask, g["point_mask"])
    plan = orc.cubic_plan(g["px"], g["py"], X, Y)
    assert (plan["simplex"] < 0).sum() == np.isnan(g["plain_out"][0]).sum() > 0
    for t, row in enumerate(g["fields"]):
        np.testing.assert_array_equal(o2.detrendedInterpolation(row.copy(), -1, "cubic"), g["mask_out"][t])
        np.testing.assert_array_equal(o2.pv, g["mask_pv"][t])
        np.testing.assert_array_equal(o2.calculateInterpolation(row.copy(), "cubic"), g["plain_out"][t])
        grad = orc.estimate_gradients(plan["tri"], row)[:, 0, :]
        v = orc.cubic_apply(plan, row, grad, X.shape)
        assert orc.parity_error(v, g["plain_out"][t], float(np.max(np.abs(row)))) < 1e-13
        # detrended: residuals on the host with the reference's coefficients, the patch is evaluated on them
        dtrend = row - (g["pz"] * o2.pv[0] + o2.pv[1])
        grad = orc.estimate_gradients(plan["tri"], dtrend)[:, 0, :]
        v = orc.cubic_apply(plan, dtrend, grad, X.shape) + o2.pv[0] * g["dem"] + o2.pv[1]
        assert orc.parity_error(v, g["mask_out"][t], float(np.max(np.abs(row)))) < 1e-13
