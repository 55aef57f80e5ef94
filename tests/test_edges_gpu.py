"""Edge cases and full-size properties (BASELINE.json shapes) on the GPU path."""
import numpy as np
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import smrf_b200.spatial as sp
    return sp


@pytest.mark.parametrize("S,ny,nx,T", [(1, 3, 5, 1), (2, 1, 1, 3), (3, 7, 9, 17), (9, 1, 130, 33), (33, 5, 13, 2)])
def test_idw_small_and_ragged_shapes(sp, S, ny, nx, T):
    from oracle import oracle as orc
    rng = np.random.default_rng(S * 100 + T)
    x = 1000.0 + 30.0 * np.arange(nx)
    y = 9000.0 - 30.0 * np.arange(ny)
    X, Y = np.meshgrid(x, y)
    dem = rng.uniform(1500, 2500, (ny, nx)).astype(np.float32).astype(np.float64)
    mx, my = rng.uniform(900, 1100 + 30 * nx, S), rng.uniform(8900 - 30 * ny, 9100, S)
    mz = rng.uniform(1400, 2600, S)
    data = rng.normal(0, 5, (T, S))
    if S > 2:
        data[0, 0] = np.nan
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    out = idw.distribute_block(data, detrend=S > 1, flag=0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(T):
            ref = o.detrendedIDW(data[t].copy(), 0) if S > 1 else o.calculateIDW(data[t].copy())
            assert parity(out[t], ref, step_scale(data[t])) <= TOL, (S, t)


def test_dk_points_grid_and_ragged_tile(sp):
    """grid_kind 1 (non-separable GridX/GridY) and a cell count that is not a multiple of the 32-cell tile."""
    from oracle import oracle as orc
    rng = np.random.default_rng(77)
    ny, nx = 7, 11                                   # 77 cells
    X = rng.uniform(0, 20000, (ny, nx))
    Y = rng.uniform(0, 20000, (ny, nx))
    dem = rng.uniform(1500, 2500, (ny, nx))
    S = 45
    mx, my, mz = rng.uniform(0, 20000, S), rng.uniform(0, 20000, S), rng.uniform(1200, 2600, S)
    data = np.abs(rng.normal(2, 2, (4, S)))
    data[2, [1, 7]] = np.nan
    cfg = {"dk_ncores": 2, "detrend_slope": 1}
    dk = sp.DK(mx, my, mz, X, Y, dem, cfg)
    o = orc.DK(mx, my, mz, X, Y, dem, cfg)
    out = dk.distribute_block(data, vmin=0.0)
    for t in range(4):
        ref = orc.set_min_max(o.calculate(data[t].copy()), 0.0, None)
        assert parity(out[t], ref, step_scale(data[t])) <= TOL, t
    dk.nan_val = np.isnan(data[2])
    o.calculate(data[2].copy())
    np.testing.assert_array_equal(dk.calculateWeights(), o.weights)


def test_grid_c5_full_size_properties(sp):
    """BASELINE config 5 shape: 4000 x 4000 cells at 10 m from ~5000 HRRR-like points.  Linear interpolation
    reproduces a linear field, NaN exactly outside the hull, clamp bounds hold, and row slabs equal the whole."""
    from smrf_b200 import synthetic as syn
    ny = nx = 4000
    x, y, dem = syn.make_dem(ny, nx, dx=10.0, noise=False)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=600.0, jitter=30.0, seed=5)
    keep = ~((px > x[-1] - 3000.0) & (py > y[0] - 2500.0))       # cut a corner: part of the DEM is outside the hull
    px, py, pz = px[keep], py[keep], pz[keep]
    X = np.broadcast_to(x[None, :], (ny, nx))
    Y = np.broadcast_to(y[:, None], (ny, nx))
    g = sp.GRID({"grid_local": False, "grid_mask": False}, px, py, X, Y, mz=pz, GridZ=dem)
    lin = np.stack([0.003 * px - 0.002 * py + 5.0, -0.001 * px + 0.004 * py - 7.0])
    out = g.distribute_block(lin, detrend=False)
    inside = g.cell_simplex.reshape(ny, nx) >= 0
    assert (~inside).sum() > 1000
    np.testing.assert_array_equal(np.isnan(out[0]), ~inside)
    want0 = 0.003 * X - 0.002 * Y + 5.0
    scale = np.abs(lin[0]).max()
    assert np.nanmax(np.abs(out[0] - want0)[inside]) <= 1e-9 * scale
    f = syn.make_source_fields(px, py, pz, 3, seed=9)
    o2 = g.distribute_block(f, detrend=True, flag=-1, vmin=-12.0, vmax=3.0)
    assert np.nanmin(o2) >= -12.0 and np.nanmax(o2) <= 3.0
    r0, r1 = 1234, 1790
    gs = sp.GRID({"grid_local": False, "grid_mask": False}, px, py, X[r0:r1], Y[r0:r1], mz=pz, GridZ=dem[r0:r1])
    np.testing.assert_array_equal(gs.distribute_block(f, detrend=True, flag=-1, vmin=-12.0, vmax=3.0), o2[:, r0:r1])


def test_idw_dk_c4_slab_properties(sp):
    """A 60-row slab of the C4 basin (10 000 columns, 200 stations): bounds, constant reproduction, and the DK
    self-check (no cell needed the exact-solver continuation after the shared walk had declared it finished)."""
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(10000, 10000, dx=30.0, noise=False)
    mx, my, mz = syn.make_stations(200, x, y, dem, seed=7)
    r0, rows = 5000, 60
    X = np.broadcast_to(x[None, :], (rows, 10000))
    Y = np.broadcast_to(y[r0:r0 + rows, None], (rows, 10000))
    Z = np.ascontiguousarray(dem[r0:r0 + rows])
    ta = syn.dropouts_iid(syn.make_series("air_temp", mz, 40, seed=3), p=0.02, seed=11)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=Z, power=2.0)
    out = idw.distribute_block(ta, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    assert np.isfinite(out).all() and out.min() >= -73.0 and out.max() <= 47.0
    const = np.full((2, 200), -3.5)
    const[1, 5::9] = np.nan
    assert np.max(np.abs(idw.distribute_block(const, detrend=False) + 3.5)) < 1e-12
    dk = sp.DK(mx, my, mz, X, Y, Z, {"detrend_slope": 1})
    pr = np.abs(syn.make_series("air_temp", mz, 3, seed=5))
    o = dk.distribute_block(np.full((1, 200), 2.25))
    assert np.max(np.abs(o - 2.25)) < 1e-9
    dk.distribute_block(pr, vmin=0.0)
    built, fallback, kmax = dk.stats()
    assert built == 1 and fallback == 0 and 2 <= kmax <= 31


def test_idw_zeros_branch_and_error_paths(sp):
    """detrendedIDW(..., zeros=mask): trend fitted on the original values, flagged stations set to zeroVal, negatives
    clipped (idw.py:96-136); and argument errors come back as exceptions with a message, never as exit()."""
    from oracle import oracle as orc
    from smrf_b200._lib import SmrfbError
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(20, 30, dx=30.0, seed=4)
    X, Y = np.meshgrid(x, y)
    mx, my, mz = syn.make_stations(12, x, y, dem, seed=2)
    row = np.abs(syn.make_series("air_temp", mz, 1, seed=8)[0]) * 0.1
    zeros = np.zeros(12, dtype=bool)
    zeros[[2, 5]] = True
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    keep = row.copy()
    v = idw.detrendedIDW(row, 1, zeros=zeros)
    np.testing.assert_array_equal(row, keep)                       # caller's row untouched (the reference mutates it)
    ref = o.detrendedIDW(keep.copy(), 1, zeros=zeros)
    assert parity(v, ref, max(step_scale(keep), 1.0)) <= TOL
    assert (v >= 0).all()
    with pytest.raises(ValueError):
        idw.distribute_block(np.zeros((2, 11)))                    # wrong station count
    with pytest.raises(SmrfbError) as e:
        sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0, device=99)
    assert "device" in str(e.value)
    with pytest.raises(SmrfbError):                                # detrend without elevations
        sp.IDW(mx, my, X, Y).distribute_block(row[None, :], detrend=True)
    big = np.linspace(0, 1, 300)
    # more stations than the tensor-core kernels hold (224): fine on a uniform mesh (near / far-field kernels, any count),
    # refused with SMRFB_ERR_LIMIT on an arbitrary point grid
    assert sp.IDW(big, big, X, Y, mz=big, GridZ=dem).kernel_info()["far_field"]
    with pytest.raises(SmrfbError) as e:
        sp.IDW(big, big, X.T.copy(), Y.T.copy(), mz=big, GridZ=dem.T.copy())
    assert e.value.code == 4
