"""GPU parity of the int8-sliced tensor-core IDW kernel (smrf_b200/csrc/idw_i8.cu; batches of >= 32 steps take it by
default, SMRFB_IDW_KERNEL=i8 forces it) against the CPU oracle, in both of its modes: denominators per step
(every step has its own NaN mask) and per distinct mask (block outages), plus the exact fp64 fallbacks
(a missing station that dominates a cell, a station exactly on a cell).  Tolerance 1e-9 relative (north_star)."""
import numpy as np
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


@pytest.fixture()
def sp_i8(monkeypatch):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    monkeypatch.setenv("SMRFB_IDW_KERNEL", "i8")          # read when a handle is created
    import smrf_b200.spatial as sp
    return sp


def _basin(ny, nx, S, T, seed=0, p=0.02, var="air_temp"):
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(ny, nx, dx=30.0, seed=seed + 100)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=seed + 7)
    data = syn.make_series(var, mz, T, seed=seed + 3)
    if p > 0:
        data = syn.dropouts_iid(data, p=p, seed=seed + 11)
    return x, y, dem, mx, my, mz, data


def _check(sp, orc, x, y, dem, mx, my, mz, data, power=2.0, detrend=True, flag=-1, vmin=-73.0, vmax=47.0):
    X, Y = np.meshgrid(x, y)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=power)
    out = idw.distribute_block(data, detrend=detrend, flag=flag, vmin=vmin, vmax=vmax)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=power)
    worst = 0.0
    for t in range(data.shape[0]):
        ref = o.detrendedIDW(data[t].copy(), flag) if detrend else o.calculateIDW(data[t].copy())
        ref = orc.set_min_max(ref, vmin, vmax)
        worst = max(worst, parity(out[t], ref, step_scale(data[t])))
    return worst, out


def test_i8_per_step_masks_vs_oracle(sp_i8):
    """200 stations (7 K blocks), 70 steps = one full chunk + a ragged one, odd cell count (ragged last tile),
    i.i.d. dropouts: every step has its own mask, so the denominators are built per step."""
    from oracle import oracle as orc
    x, y, dem, mx, my, mz, data = _basin(37, 53, 200, 70, seed=1)
    rng = np.random.default_rng(5)
    data[10, rng.random(200) < 0.4] = np.nan
    data[11, 5:] = np.nan                                   # 5 valid stations only
    worst, _ = _check(sp_i8, orc, x, y, dem, mx, my, mz, data)
    assert worst <= TOL, worst


@pytest.mark.parametrize("S", [150, 33, 224])
def test_i8_mask_mode_vs_oracle(sp_i8, S):
    """Block outages: a handful of distinct masks, denominators once per tile and mask.  Station counts with
    5, 2 and 7 K blocks."""
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem, mx, my, mz, data = _basin(29, 41, S, 100, seed=2, p=0.0)
    data, nm = syn.dropouts_block(data, max_masks=7, per_year=300.0, mean_len=20.0)
    assert 1 < nm <= 7 and np.isnan(data).any()
    worst, _ = _check(sp_i8, orc, x, y, dem, mx, my, mz, data, flag=0)
    assert worst <= TOL, worst
    worst, _ = _check(sp_i8, orc, x, y, dem, mx, my, mz, data, detrend=False, vmin=-np.inf, vmax=np.inf)
    assert worst <= TOL, worst


def test_i8_power_not_two(sp_i8):
    from oracle import oracle as orc
    x, y, dem, mx, my, mz, data = _basin(20, 24, 40, 50, seed=3)
    worst, _ = _check(sp_i8, orc, x, y, dem, mx, my, mz, data, power=1.5, flag=1)
    assert worst <= TOL, worst


@pytest.mark.parametrize("mode", ["per_step", "masks"])
def test_i8_dominant_station_missing_and_coincident(sp_i8, mode):
    """A station 1 mm from a cell centre holds all but 1e-13 of that cell's weight; when it drops out the
    fixed-point digits of the others are exhausted and the kernel must take its exact fp64 path.
    A station exactly on a cell centre has an infinite weight (H6: NaN where its residual is not zero)."""
    from oracle import oracle as orc
    x, y, dem, mx, my, mz, data = _basin(16, 16, 40, 60, seed=2, p=0.0)
    mx[3], my[3] = x[4] + 1e-3, y[9]
    mx[8], my[8] = x[11], y[2] - 2e-4
    mx[20], my[20] = x[7], y[7]                            # coincident
    if mode == "per_step":
        rng = np.random.default_rng(1)
        data[rng.random(data.shape) < 0.05] = np.nan
    data[1:20, 3] = np.nan
    data[10:30, 8] = np.nan
    data[25:40, 20] = np.nan
    worst, out = _check(sp_i8, orc, x, y, dem, mx, my, mz, data, flag=0, vmin=-np.inf, vmax=np.inf)
    assert worst <= TOL, worst
    assert np.isnan(out[0][7, 7]) and not np.isnan(out[30][7, 7])


def test_i8_large_grid_properties(sp_i8, monkeypatch):
    """Grid too big for the oracle: agreement with the fp64 tensor-core kernel, constants, row slabs, window calls."""
    x, y, dem, mx, my, mz, data = _basin(300, 500, 200, 96, seed=3)
    X, Y = np.meshgrid(x, y)
    idw = sp_i8.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    full = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    monkeypatch.setenv("SMRFB_IDW_KERNEL", "dmma")
    ref = sp_i8.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0).distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    monkeypatch.setenv("SMRFB_IDW_KERNEL", "i8")
    scale = np.nanmax(np.abs(data), axis=1)[:, None, None]
    assert np.max(np.abs(full - ref) / scale) <= 2e-10          # two different summation orders, both inside 1e-9 of numpy
    const = np.full((40, 200), 7.25)
    const[1::3, ::7] = np.nan
    out = idw.distribute_block(const, detrend=False)
    assert np.max(np.abs(out - 7.25)) < 1e-12
    r0, r1 = 123, 251
    slab = sp_i8.IDW(mx, my, X[r0:r1], Y[r0:r1], mz=mz, GridZ=dem[r0:r1], power=2.0)
    np.testing.assert_array_equal(slab.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0), full[:, r0:r1])
