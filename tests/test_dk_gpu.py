"""GPU parity of detrended kriging (through the C-ABI): active sets and weights against the
reference-generated golden vectors and the CPU oracle (bit-exact active set; weights are the
reference's own last-iteration arithmetic, so they are compared bit for bit too), distributed
fields at 1e-9."""
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


def test_krige_kat_call_grid(sp, golden):
    g = golden("krige_kat")
    w = np.zeros_like(g["dgrid"])
    sp.detrended_kriging.call_grid(g["ad"], g["dgrid"], g["sz"], w, 2)
    np.testing.assert_array_equal(w != 0, g["weights"] != 0)        # active set, bit-exact
    np.testing.assert_array_equal(w, g["weights"])                  # and the weights themselves


@pytest.mark.parametrize("nsta,ncell,seed", [(1, 5, 9), (2, 40, 0), (5, 200, 1), (13, 333, 2), (40, 500, 3),
                                             (100, 300, 4), (200, 96, 5), (231, 48, 6)])   # 231 > 200: the inverse leaves shared memory
def test_krige_grid_vs_oracle(sp, nsta, ncell, seed):
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    sx, sy = rng.uniform(0, 30000, nsta), rng.uniform(0, 30000, nsta)
    sz = rng.uniform(900, 3000, nsta)
    if nsta > 4:
        sz[1] = -5.0                  # elevation <= 0 is never dropped (krige.c:172)
        sz[3] = sz[2]                 # elevation tie: lowest index wins
    cx, cy = rng.uniform(-2000, 32000, ncell), rng.uniform(-2000, 32000, ncell)
    ad = orc.station_distances(sx, sy)
    dg = np.sqrt((cx[:, None] - sx[None, :]) ** 2 + (cy[:, None] - sy[None, :]) ** 2)
    w_ref = np.zeros_like(dg)
    orc.call_grid(ad, dg, sz, w_ref, 0)
    w = np.zeros_like(dg)
    sp.detrended_kriging.call_grid(ad, dg, sz, w, 1)
    np.testing.assert_array_equal(w != 0, w_ref != 0)
    np.testing.assert_array_equal(w, w_ref)


def test_dk_synth_golden(sp, golden):
    g = golden("dk_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    dk = sp.DK(g["mx"], g["my"], g["mz"], X, Y, g["dem"], {"dk_ncores": 2, "detrend_slope": 1})
    out = dk.distribute_block(g["data"], vmin=0.0)
    for t in range(len(g["data"])):
        assert parity(out[t], g["out"][t], step_scale(g["data"][t])) <= TOL, t
    # per-step reference API, weights per mask
    for t in (0, 2, 4):
        v = dk.calculate(g["data"][t].copy())
        np.testing.assert_array_equal(dk.nan_val, g["nan_mask"][t])
        w = dk.calculateWeights()
        full = np.zeros(X.shape + (len(g["mx"]),))
        full[:, :, ~dk.nan_val] = w
        np.testing.assert_array_equal(full, g["weights"][t])
        from smrf_b200.utils import set_min_max
        assert parity(set_min_max(v, 0.0, None), g["out"][t], step_scale(g["data"][t])) <= TOL
    built, fallback, kmax = dk.stats()
    assert built == 3            # three distinct NaN masks in the fixture


def test_dk_rme_golden(sp, golden):
    g = golden("dk_rme")
    X, Y = np.meshgrid(g["x"], g["y"])
    dk = sp.DK(g["mx"], g["my"], g["mz"], X, Y, g["dem"], {"dk_ncores": 2, "detrend_slope": 1})
    out = dk.distribute_block(g["data"], vmin=0.0)
    for t in range(len(g["data"])):
        assert parity(out[t], g["out"][t], step_scale(g["data"][t])) <= TOL, t


def test_dk_200_stations_vs_oracle(sp):
    """C4-shaped slice: 200 stations, a DEM patch, block outages -> a few masks."""
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(4000, 4000, dx=30.0, noise=False)
    mx, my, mz = syn.make_stations(200, x, y, dem, seed=7)
    r0, c0 = 1500, 2100
    xs, ys, dems = x[c0:c0 + 48], y[r0:r0 + 20], dem[r0:r0 + 20, c0:c0 + 48]
    X, Y = np.meshgrid(xs, ys)
    data = syn.make_series("precip", mz, 60, seed=5)
    data = data[data.sum(axis=1) > 0][:5]
    data[2:, [3, 50, 117]] = np.nan
    data[4, 199] = np.nan
    cfg = {"dk_ncores": 8, "detrend_slope": 1}
    dk = sp.DK(mx, my, mz, X, Y, dems, cfg)
    out = dk.distribute_block(data, vmin=0.0)
    o = orc.DK(mx, my, mz, X, Y, dems, cfg)
    for t in range(len(data)):
        ref = orc.set_min_max(o.calculate(data[t].copy()), 0.0, None)
        assert parity(out[t], ref, step_scale(data[t])) <= TOL, t
        if t in (0, 2):
            dk.nan_val = np.isnan(data[t])
            w = dk.calculateWeights()
            np.testing.assert_array_equal(w != 0, o.weights != 0)       # bit-exact active set
            np.testing.assert_array_equal(w, o.weights)
    np.testing.assert_allclose(dk.pv_block[-1], o.pv, rtol=1e-9, atol=1e-12)


def test_dk_properties(sp):
    """Weights sum to 1, are non-negative, and a constant field comes back constant; row slabs
    (the multi-GPU partition) equal the whole."""
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(150, 170, dx=60.0, seed=12)
    mx, my, mz = syn.make_stations(60, x, y, dem, seed=3)
    X, Y = np.meshgrid(x, y)
    cfg = {"dk_ncores": 1, "detrend_slope": 0}
    dk = sp.DK(mx, my, mz, X, Y, dem, cfg)
    dk.nan_val = np.zeros(60, dtype=bool)
    w = dk.calculateWeights()
    assert (w >= 0).all()
    np.testing.assert_allclose(w.sum(axis=2), 1.0, rtol=0, atol=1e-9)
    out = dk.distribute_block(np.full((2, 60), 3.5))
    assert np.max(np.abs(out - 3.5)) < 1e-9
    data = syn.make_series("air_temp", mz, 6, seed=1)
    full = dk.distribute_block(data)
    slab = sp.DK(mx, my, mz, X[40:97], Y[40:97], dem[40:97], cfg)
    # the tensor-core steady-state kernel sums a cell's terms in the order of its warp's station union, which depends on
    # where the slab starts: equal to a few ulp, not bit for bit
    np.testing.assert_allclose(slab.distribute_block(data), full[:, 40:97], rtol=1e-12, atol=1e-12)
    # exactness at a station: a cell coincident with station 5 reproduces its value
    mx2, my2 = mx.copy(), my.copy()
    mx2[5], my2[5] = x[33], y[21]
    dk2 = sp.DK(mx2, my2, mz, X, Y, dem, cfg)
    o = dk2.distribute_block(data)
    pv = dk2.pv_block
    np.testing.assert_allclose(o[:, 21, 33], data[:, 5] - (mz[5] * pv[:, 0] + pv[:, 1]) + pv[:, 0] * dem[21, 33] + pv[:, 1],
                               rtol=1e-9, atol=1e-9)


def test_dk_singular_is_an_error_not_exit(sp):
    from smrf_b200._lib import SmrfbError
    X, Y = np.meshgrid(np.arange(6.0) * 10, np.arange(5.0) * 10)
    mx = np.array([3.0, 3.0, 40.0])
    my = np.array([4.0, 4.0, 30.0])                    # two co-located stations (krige.c:161-164 exit()s)
    dk = sp.DK(mx, my, np.array([1.0, 2.0, 3.0]), X, Y, np.ones((5, 6)), {"detrend_slope": 0})
    with pytest.raises(SmrfbError) as e:
        dk.distribute_block(np.array([[1.0, 2.0, 3.0]]))
    assert e.value.code == 3


def test_dk_c4_scattered_cells_bit_exact(sp):
    """2048 cells scattered over the whole 300 km C4 domain (every cell walks its own elimination path,
    ~190 drops each): active sets and weights bit-identical to the reference algorithm."""
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(10000, 10000, dx=30.0, noise=False)
    mx, my, mz = syn.make_stations(200, x, y, dem, seed=7)
    rng = np.random.default_rng(123)
    ii = rng.integers(0, 10000, size=(32, 64))
    jj = rng.integers(0, 10000, size=(32, 64))
    X, Y, Z = x[jj], y[ii], dem[ii, jj]
    cfg = {"dk_ncores": 16, "detrend_slope": 1}
    dk = sp.DK(mx, my, mz, X, Y, Z, cfg)
    dk.nan_val = np.zeros(200, dtype=bool)
    dk.nan_val[[11, 42, 190]] = True
    w = dk.calculateWeights()
    o = orc.DK(mx, my, mz, X, Y, Z, cfg)
    o.nan_val = dk.nan_val.copy()
    o.calculateWeights()
    np.testing.assert_array_equal(w != 0, o.weights != 0)
    np.testing.assert_array_equal(w, o.weights)
    assert dk.stats()[1] == 0
