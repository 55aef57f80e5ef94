"""GPU parity of the IDW path (through the C-ABI) against the reference-generated golden vectors
and the CPU oracle.  Tolerance: 1e-9 relative (north_star), metric of SURVEY.md H7; NaN patterns
and the station mask must match exactly."""
import warnings

import numpy as np
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default", "dmma"])
def sp(request, monkeypatch):
    """default: near field + interpolated far field on uniform meshes (idw_far.cu), else the tensor-core kernels;
    dmma: the dense fp64 tensor-core kernel for every call (SMRFB_IDW_KERNEL is read when a handle is created)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    if request.param == "dmma":
        monkeypatch.setenv("SMRFB_IDW_KERNEL", "dmma")
    else:
        monkeypatch.delenv("SMRFB_IDW_KERNEL", raising=False)
    import smrf_b200.spatial as sp
    return sp


def test_idw_rme_golden(sp, golden):
    g = golden("idw_rme")
    X, Y = np.meshgrid(g["x"], g["y"])
    idw = sp.IDW(g["mx"], g["my"], X, Y, mz=g["mz"], GridZ=g["dem"], power=2.0)
    for name, lo, hi, det in (("air_temp", -73.0, 47.0, True), ("vapor_pressure", 20.0, 5000.0, True),
                              ("cloud_factor", 0.0, 1.0, False)):
        data = g[name]
        out = idw.distribute_block(data, detrend=det, flag=-1, vmin=lo, vmax=hi)
        for t in range(len(data)):
            assert parity(out[t], g["out_" + name][t], step_scale(data[t])) <= TOL
        if name == "air_temp":
            # coefficients agree wherever the fit is not degenerate (equal station values give a slope of
            # +-1e-19 in np.polyfit and exactly 0 here; the sign constraint may then zero one and not the
            # other -- the distributed field is the same either way, which the loop above checked)
            ref_pv = g["pv_air_temp"]
            generic = np.abs(data[:, 0] - data[:, 1]) > 1e-9
            np.testing.assert_allclose(idw.pv_block[generic], ref_pv[generic], rtol=1e-9, atol=1e-9)
    # the per-step reference API gives the same thing, and leaves the caller's row untouched
    row = g["air_temp"][3].copy()
    v = idw.detrendedIDW(row, -1)
    np.testing.assert_array_equal(row, g["air_temp"][3])
    from smrf_b200.utils import set_min_max
    v = set_min_max(v, -73.0, 47.0)
    assert parity(v, g["out_air_temp"][3], step_scale(row)) <= TOL
    assert v.flags["C_CONTIGUOUS"] and v.dtype == np.float64 and v.shape == X.shape


@pytest.mark.parametrize("power", [2.0, 1.5])
def test_idw_synth_golden(sp, golden, power):
    g = golden("idw_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    idw = sp.IDW(g["mx"], g["my"], X, Y, mz=g["mz"], GridZ=g["dem"], power=power)
    data = g["data"]
    for flag in (-1, 0, 1):
        out = idw.distribute_block(data, detrend=True, flag=flag)
        for t in range(len(data)):
            assert parity(out[t], g["out_p%g_f%d" % (power, flag)][t], step_scale(data[t])) <= TOL, (flag, t)
    out = idw.distribute_block(data, detrend=False)
    for t in range(len(data)):
        assert parity(out[t], g["out_p%g_plain" % power][t], step_scale(data[t])) <= TOL
        v = idw.calculateIDW(data[t])
        np.testing.assert_array_equal(v, out[t])          # per-step call == batched call, bit for bit


def test_idw_zero_distance_quirk(sp, golden):
    g = golden("idw_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    idw = sp.IDW(g["mx0"], g["my0"], X, Y, mz=g["mz"], GridZ=g["dem"], power=2.0)
    out = idw.distribute_block(g["data"][:4], detrend=True, flag=0)
    for t in range(4):
        assert parity(out[t], g["out_zero_dist"][t], step_scale(g["data"][t])) <= TOL
    assert np.isnan(out[1][7, 5]) and not np.isnan(out[0][7, 5])


def _basin(ny, nx, S, T, seed=0, p=0.02):
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(ny, nx, dx=30.0, seed=seed + 100)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=seed + 7)
    data = syn.dropouts_iid(syn.make_series("air_temp", mz, T, seed=seed + 3), p=p, seed=seed + 11)
    return x, y, dem, mx, my, mz, data


def test_idw_200_stations_vs_oracle(sp):
    from oracle import oracle as orc
    x, y, dem, mx, my, mz, data = _basin(37, 53, 200, 70, seed=1)       # odd cell count, T not a multiple of 32
    rng = np.random.default_rng(5)
    data[10, rng.random(200) < 0.4] = np.nan                             # > 13 missing: dense-denominator path
    data[11, 5:] = np.nan                                                # 5 valid stations only
    X, Y = np.meshgrid(x, y)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    out = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    worst = 0.0
    for t in range(data.shape[0]):
        ref = orc.set_min_max(o.detrendedIDW(data[t].copy(), -1), -73.0, 47.0)
        worst = max(worst, parity(out[t], ref, step_scale(data[t])))
        np.testing.assert_allclose(idw.pv_block[t], o.pv, rtol=1e-9, atol=1e-12)
    assert worst <= TOL, worst
    # grid_kind 1 (arbitrary cell coordinates) gives the same numbers as the mesh form
    Xp = X.copy()
    Xp[0, 0] += 0.0                                                      # still a mesh -> force points via transpose trick
    idw_pts = sp.IDW(mx, my, X.T.copy(), Y.T.copy(), mz=mz, GridZ=dem.T.copy(), power=2.0)
    out_pts = idw_pts.distribute_block(data[:3], detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    out_3 = idw.distribute_block(data[:3], detrend=True, flag=-1, vmin=-73.0, vmax=47.0)   # same batch length: same kernel
    # (bit-identical when both take the same kernel; the mesh form may take the near / far-field kernels instead)
    assert max(parity(out_pts[t].T, out_3[t], step_scale(data[t])) for t in range(3)) <= 1e-11
    for t in range(3):                                                                    # short and long batches agree
        assert parity(out_3[t], out[t], step_scale(data[t])) <= TOL


def test_idw_missing_station_next_to_cell(sp):
    """A missing station 1 mm from a cell centre dominates the all-station weight sum by ~1e13; the
    masked denominator must not lose those digits (double-double path)."""
    from oracle import oracle as orc
    x, y, dem, mx, my, mz, data = _basin(16, 16, 24, 6, seed=2, p=0.0)
    mx[3], my[3] = x[4] + 1e-3, y[9]
    mx[8], my[8] = x[11], y[2] - 2e-4
    data[1, 3] = np.nan
    data[2, [3, 8]] = np.nan
    data[4, 8] = np.nan
    X, Y = np.meshgrid(x, y)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    out = idw.distribute_block(data, detrend=True, flag=0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    for t in range(data.shape[0]):
        ref = o.detrendedIDW(data[t].copy(), 0)
        assert parity(out[t], ref, step_scale(data[t])) <= TOL, t


def test_idw_properties_large(sp):
    """Size-independent properties on a grid too big for the oracle: constants are reproduced,
    row slabs (the multi-GPU partition) equal the whole, masked denominators are consistent."""
    x, y, dem, mx, my, mz, data = _basin(600, 700, 200, 40, seed=3)
    X, Y = np.meshgrid(x, y)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    const = np.full((3, 200), 7.25)
    const[1, ::7] = np.nan
    out = idw.distribute_block(const, detrend=False)
    assert np.max(np.abs(out - 7.25)) < 1e-12          # design bound: D - sum(missing) may lose up to ~7 bits => < 1e-13 relative
    full = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    assert np.isfinite(full).all() and full.min() >= -73.0 and full.max() <= 47.0
    r0, r1 = 123, 391
    slab = sp.IDW(mx, my, X[r0:r1], Y[r0:r1], mz=mz, GridZ=dem[r0:r1], power=2.0)
    out_slab = slab.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    # bit-identical on the dense kernels; the near / far-field kernels cut a sub-grid into different patches (row windows of
    # ONE handle are bit-identical there: test_idw_far_gpu.py::test_far_row_windows_and_float32)
    assert np.max(np.abs(out_slab - full[:, r0:r1])) <= 1e-10 * np.nanmax(np.abs(data))
    # linearity in the data for a fixed mask (no detrend, no clamp): IDW(a*d1 + d2) = a*IDW(d1) + IDW(d2)
    d1, d2 = data[:4].copy(), data[4:8].copy()
    m = np.isnan(d1) | np.isnan(d2)
    d1[m] = np.nan
    d2[m] = np.nan
    lhs = idw.distribute_block(2.5 * d1 + d2, detrend=False)
    rhs = 2.5 * idw.distribute_block(d1, detrend=False) + idw.distribute_block(d2, detrend=False)
    assert np.max(np.abs(lhs - rhs)) <= 1e-9 * np.nanmax(np.abs(data))


def test_idw_all_nan_row_raises(sp):
    x, y, dem, mx, my, mz, data = _basin(8, 8, 6, 3, seed=4, p=0.0)
    X, Y = np.meshgrid(x, y)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem)
    data[1, :] = np.nan
    from smrf_b200._lib import SmrfbError
    with pytest.raises(SmrfbError):
        idw.distribute_block(data, detrend=True, flag=0)


def test_idw_small_station_count_both_kernels(sp, golden, monkeypatch):
    """S <= 16 normally takes the register-weight kernel; the tensor-core kernel must give the same field."""
    g = golden("idw_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    from oracle import oracle as orc
    sel = np.arange(0, 20, 2)                                  # 10 of the 20 stations
    idw = sp.IDW(g["mx"][sel], g["my"][sel], X, Y, mz=g["mz"][sel], GridZ=g["dem"], power=2.0)
    o = orc.IDW(g["mx"][sel], g["my"][sel], X, Y, mz=g["mz"][sel], GridZ=g["dem"], power=2.0)
    data = g["data"][:4, sel]
    a = idw.distribute_block(data, detrend=True, flag=-1, vmin=-10.0, vmax=30.0)
    monkeypatch.setenv("SMRFB_IDW_FORCE_MMA", "1")          # read when the handle is created
    idw_mma = sp.IDW(g["mx"][sel], g["my"][sel], X, Y, mz=g["mz"][sel], GridZ=g["dem"], power=2.0)
    monkeypatch.delenv("SMRFB_IDW_FORCE_MMA")
    b = idw_mma.distribute_block(data, detrend=True, flag=-1, vmin=-10.0, vmax=30.0)
    for t in range(len(data)):
        want = np.clip(o.detrendedIDW(data[t].copy(), -1), -10.0, 30.0)
        assert parity(a[t], want, step_scale(data[t])) <= TOL
        assert parity(b[t], want, step_scale(data[t])) <= TOL
