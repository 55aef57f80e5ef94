"""Parity of the near field + interpolated far field IDW kernels (smrf_b200/csrc/idw_far.cu, the default on uniform
meshes) against the CPU oracle (smrf/spatial/idw.py:53-144 restated in oracle/oracle.py) on the HEADLINE geometry of
BASELINE.json configs[3]: a window of a 30 m mesh inside a 300 km station network, so that most stations are FAR from
every patch and enter only through the Chebyshev interpolant, a few are near, one sits 12 m from a cell centre and one
exactly on a cell centre (the reference's zero-distance case, idw.py:69).  Tolerance 1e-9 relative (north_star); the
kernel's a-priori bound is 2e-10."""
import ctypes

import numpy as np
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu

CLAMP = {"air_temp": (-73.0, 47.0), "vapor_pressure": (20.0, 5000.0)}


@pytest.fixture()
def sp(monkeypatch):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    monkeypatch.delenv("SMRFB_IDW_KERNEL", raising=False)
    import smrf_b200.spatial as sp
    return sp


def window_basin(ny, nx, S, seed=5, x_off=141000.0, y_off=97000.0, dx=30.0, span=300e3, extra=True):
    """A [ny, nx] window of the C4 mesh at (x_off, y_off) inside a span x span station network."""
    rng = np.random.default_rng(seed)
    x = 500000.0 + x_off + dx * np.arange(nx)
    y = 4800000.0 - y_off - dx * np.arange(ny)
    X, Y = np.meshgrid(x, y)
    dem = (1800.0 + 700.0 * np.sin(2 * np.pi * X / 41000.0) * np.cos(2 * np.pi * Y / 57000.0)
           + 120.0 * np.sin(2 * np.pi * X / 3100.0 + 1.0) * np.sin(2 * np.pi * Y / 2300.0)).astype(np.float32).astype(np.float64)
    mx = 500000.0 + rng.uniform(0, span, S)
    my = 4800000.0 - rng.uniform(0, span, S)
    if extra:
        # inside the window: 12 m off a cell centre, exactly on a cell centre, and two more within a few km
        mx[0], my[0] = x[nx // 3] + 12.0, y[ny // 2] - 3.0
        mx[1], my[1] = x[2 * nx // 3], y[ny // 4]
        mx[2], my[2] = x[0] - 2500.0, y[ny - 1] - 900.0
        mx[3], my[3] = x[nx - 1] + 7000.0, y[0] + 4000.0
    mz = 1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0) + rng.normal(0, 5.0, S)
    return x, y, X, Y, dem, mx, my, mz


def series(var, mz, T, drop):
    from smrf_b200 import synthetic as syn
    data = syn.make_series(var, mz, T, seed=3 if var == "air_temp" else 4)
    if drop == "iid":
        data = syn.dropouts_iid(data, p=0.03, seed=11)
        data[T - 1, : len(mz) // 3] = np.nan              # a third of the network down, incl. the stations in the window
        data[T - 2, 1] = np.nan                           # the coincident station missing
    else:
        data, nm = syn.dropouts_block(data, max_masks=5, per_year=400.0, mean_len=30.0, seed=13)
        assert nm > 1
    return data


@pytest.mark.parametrize("var,drop,power", [("air_temp", "iid", 2.0), ("vapor_pressure", "block", 2.0),
                                            ("air_temp", "block", 1.5), ("vapor_pressure", "iid", 3.0)])
def test_far_vs_oracle(sp, var, drop, power):
    from oracle import oracle as orc
    x, y, X, Y, dem, mx, my, mz = window_basin(203, 301, 200)
    T = 40
    data = series(var, mz, T, drop)
    lo, hi = CLAMP[var]
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=power)
    info = idw.kernel_info()
    # powers <= 2: 10 x 10 nodes from theta = 6, or -- dense networks like this window -- 12 x 12 nodes from theta = 4
    assert info["far_field"] and info["max_near"] >= 3
    assert (info["nodes"], info["theta"]) in (((10, 6.0), (12, 4.0)) if power <= 2 else ((12, 8.0),))
    out = idw.distribute_block(data, detrend=True, flag=-1, vmin=lo, vmax=hi)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=power)
    worst = 0.0
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                   # the coincident station: divide by zero, as in the reference
        for t in range(T):
            ref = orc.set_min_max(o.detrendedIDW(data[t].copy(), -1), lo, hi)
            worst = max(worst, parity(out[t], ref, step_scale(data[t])))
        assert worst <= TOL, worst
        out = idw.distribute_block(data[:3], detrend=False)             # short batch, undetrended, unclamped
        worst2 = max(parity(out[t], o.calculateIDW(data[t].copy()), step_scale(data[t])) for t in range(3))
    assert worst2 <= TOL, worst2
    print("far field %s/%s/p=%.1f: %.2e detrended, %.2e plain" % (var, drop, power, worst, worst2))


def test_far_ten_nodes_on_a_dense_network(sp, monkeypatch):
    """SMRFB_IDW_FAR_NN10 keeps the 10 x 10 / theta = 6 plan where the library would pick 12 x 12 / theta = 4: same contract."""
    from oracle import oracle as orc
    monkeypatch.setenv("SMRFB_IDW_FAR_NN10", "1")
    x, y, X, Y, dem, mx, my, mz = window_basin(140, 260, 120, seed=4)
    data = series("air_temp", mz, 12, "iid")
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    info = idw.kernel_info()
    assert info["far_field"] and (info["nodes"], info["theta"]) == (10, 6.0)
    out = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        worst = max(parity(out[t], orc.set_min_max(o.detrendedIDW(data[t].copy(), -1), -73.0, 47.0), step_scale(data[t])) for t in range(12))
    assert worst <= TOL, worst


def test_far_coincident_station_semantics(sp):
    """A valid station exactly on a cell centre makes that cell NaN (inf / inf), a missing one does not (idw.py:69-94)."""
    from oracle import oracle as orc
    x, y, X, Y, dem, mx, my, mz = window_basin(40, 150, 60, seed=9)
    data = series("air_temp", mz, 8, "iid")
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    assert idw.kernel_info()["far_field"]
    out = idw.distribute_block(data, detrend=True, flag=0)
    ci, cj = 40 // 4, 2 * 150 // 3
    assert np.isnan(out[0, ci, cj]) and not np.isnan(out[6, ci, cj]) and not np.isnan(out[7, ci, cj])
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(8):
            assert parity(out[t], o.detrendedIDW(data[t].copy(), 0), step_scale(data[t])) <= TOL


def test_far_row_windows_and_float32(sp):
    """Row windows that start and end anywhere (not on 8-row bands or 128-row patches) reproduce the whole-grid
    result bit for bit; float32 output is the rounded float64 value."""
    import torch
    from smrf_b200 import _lib
    x, y, X, Y, dem, mx, my, mz = window_basin(300, 260, 120, seed=3)
    T = 20
    data = series("air_temp", mz, T, "iid")
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    full = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    L = _lib.lib()
    d_data = torch.from_numpy(data).cuda()
    for r0, r1 in ((0, 300), (37, 150), (128, 136), (255, 300), (131, 132)):
        d_out = torch.full((T, (r1 - r0) * 260), 7.0, dtype=torch.float64, device="cuda")
        _lib.check(L.smrfb_idw_distribute_window_dev(idw._h, d_data.data_ptr(), T, 1, -1, -73.0, 47.0, None, None,
                                                     r0 * 260, (r1 - r0) * 260, d_out.data_ptr(), None))
        _lib.check(L.smrfb_synchronize(idw._h))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_out.cpu().numpy().reshape(T, r1 - r0, 260), full[:, r0:r1])
    f32 = idw.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0, out_dtype=np.float32)
    assert f32.dtype == np.float32
    np.testing.assert_array_equal(f32, full.astype(np.float32))


def test_far_more_stations_than_the_tensor_core_kernels_hold(sp):
    from oracle import oracle as orc
    x, y, X, Y, dem, mx, my, mz = window_basin(64, 200, 300, seed=21)
    data = series("vapor_pressure", mz, 6, "iid")
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    assert idw.kernel_info()["far_field"]
    out = idw.distribute_block(data, detrend=True, flag=-1, vmin=20.0, vmax=5000.0)
    o = orc.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        worst = max(parity(out[t], orc.set_min_max(o.detrendedIDW(data[t].copy(), -1), 20.0, 5000.0), step_scale(data[t]))
                    for t in range(6))
    assert worst <= TOL, worst


def test_far_vs_fp64_tensor_core_kernel_large_grid(sp, monkeypatch):
    """1100 x 1300 cells (9 x 11 patches, ragged edges): the far-field path against the dense fp64 tensor-core kernel."""
    x, y, X, Y, dem, mx, my, mz = window_basin(1100, 1300, 200, seed=8, extra=False)
    T = 33
    data = series("air_temp", mz, T, "iid")
    a = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    assert a.kernel_info()["far_field"]
    va = a.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    monkeypatch.setenv("SMRFB_IDW_KERNEL", "dmma")
    b = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    assert not b.kernel_info()["far_field"]
    vb = b.distribute_block(data, detrend=True, flag=-1, vmin=-73.0, vmax=47.0)
    worst = max(parity(va[t], vb[t], step_scale(data[t])) for t in range(T))
    assert worst <= TOL, worst
    print("far field vs dense fp64 on 1100 x 1300: %.2e" % worst)


def test_far_not_used_on_irregular_meshes(sp):
    x, y, X, Y, dem, mx, my, mz = window_basin(30, 50, 40, seed=2)
    xs = x.copy()
    xs[10:] += 3.0                                         # a non-uniform axis
    Xs, Ys = np.meshgrid(xs, y)
    assert not sp.IDW(mx, my, Xs, Ys, mz=mz, GridZ=dem, power=2.0).kernel_info()["far_field"]
    assert not sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=5.0).kernel_info()["far_field"]
    assert sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0).kernel_info()["far_field"]
