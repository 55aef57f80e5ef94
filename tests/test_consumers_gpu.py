"""GPU parity of the components either side of the interpolators: dew point (stand-alone and as a consumer hooked onto a
distribute call), wind azimuth, WindNinja-style weights interpolation, the variable wrappers (ta, vp, cf, ppt, Wind) driven
like smrf.framework's loop, the block façade and the float32 sink.  Oracle = oracle/ (CPU restatements pinned to the
reference, tests/test_consumers_cpu.py and tests/test_oracle.py)."""
import os

import numpy as np
import pandas as pd
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _dew_close(got, ref, tol):
    """The device's pow / log differ from glibc's in the last bit, so an iterate of the Brent search may differ by ~1e-13
    before the float truncation: equal to one float ulp almost everywhere, and never further apart than the routine's own
    tolerance (a branch of the search may flip on a borderline test)."""
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    d = np.abs(got[ok] - ref[ok])
    assert np.mean(d <= 8e-6) >= 0.999, np.mean(d <= 8e-6)
    assert d.max() <= 2.0 * tol + 1e-5, d.max()


def test_dew_point_standalone_and_golden(gpu, golden):
    from oracle import oracle as orc
    from smrf_b200 import envphys_c
    g = golden("dewpt")
    for tol in (0.01, 0.001):
        dpt = np.zeros_like(g["vp"])
        envphys_c.cdewpt(g["vp"], dpt, tol, 4)
        _dew_close(dpt, g["dpt_tol%g" % tol], tol)
    vp = np.random.default_rng(5).uniform(20.0, 5000.0, (37, 53))
    vp[3, 4] = np.nan
    dpt = np.empty_like(vp)
    envphys_c.cdewpt(vp, dpt, 0.01, 1)
    _dew_close(dpt, orc.dew_point(vp, 0.01), 0.01)


def test_wet_bulb_standalone_and_golden(gpu, golden):
    """smrfb_wet_bulb (envphys_c.cwbt, iwbt.c:84-134) against the reference-generated vectors: the device's pow / log differ from
    glibc's in the last bit, so the iterates agree to ~1e-12; a borderline `step > tol` test may add or drop one Newton step,
    which moves the result by less than the routine's tolerance."""
    from smrf_b200 import envphys_c
    from smrf_b200._lib import SmrfbError
    g = golden("wetbulb")
    for tol in (0.01, 0.05):
        tw = np.zeros_like(g["ta"])
        envphys_c.cwbt(g["ta"], g["td"], g["z"], tw, tol, 4)
        d = np.abs(tw - g["tw_tol%g" % tol])
        assert np.mean(d <= 1e-9) >= 0.999, np.mean(d <= 1e-9)
        assert d.max() <= tol, d.max()
    ta = np.random.default_rng(2).uniform(-20.0, 25.0, (31, 47))
    tw = np.empty_like(ta)
    envphys_c.cwbt(ta, ta - 3.0, np.full_like(ta, 1800.0), tw, 0.01)
    assert np.all(tw <= ta + 1e-9) and np.all(tw >= ta - 3.0 - 1e-9)
    with pytest.raises(SmrfbError):                               # below 0 K: the reference prints and exits
        envphys_c.cwbt(np.array([-300.0]), np.array([-300.0]), np.array([0.0]), np.empty(1), 0.01)
    with pytest.raises(SmrfbError):                               # tolerance 0 can only end in the reference's exit(-1)
        envphys_c.cwbt(np.array([5.0]), np.array([1.0]), np.array([0.0]), np.empty(1), 0.0)


def test_dew_point_consumer_on_idw_dk_grid(gpu):
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    import smrf_b200.spatial as sp
    x, y, dem = syn.make_dem(60, 90, dx=30.0, seed=4)
    X, Y = np.meshgrid(x, y)
    mx, my, mz = syn.make_stations(24, x, y, dem, seed=9)
    vp = syn.dropouts_iid(syn.make_series("vapor_pressure", mz, 5, seed=4), p=0.1, seed=2)
    idw = sp.IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    plain = idw.distribute_block(vp, detrend=True, flag=-1, vmin=20.0, vmax=5000.0)
    out, dpt = idw.dew_point_of(lambda: idw.distribute_block(vp, detrend=True, flag=-1, vmin=20.0, vmax=5000.0), 5, 0.01)
    np.testing.assert_array_equal(out, plain)                       # the hook does not disturb the primary output
    _dew_close(dpt, orc.dew_point(out, 0.01), 0.01)
    again = idw.distribute_block(vp, detrend=True, flag=-1, vmin=20.0, vmax=5000.0)   # hook removed again
    np.testing.assert_array_equal(again, plain)
    dk = sp.DK(mx, my, mz, X, Y, dem, {"dk_ncores": 1, "detrend_slope": 0})
    out, dpt = dk.dew_point_of(lambda: dk.distribute_block(vp, vmin=20.0, vmax=5000.0), 5, 0.001)
    _dew_close(dpt, orc.dew_point(out, 0.001), 0.001)
    out32, dpt32 = idw.dew_point_of(lambda: idw.distribute_block(vp, detrend=True, flag=-1, vmin=20.0, vmax=5000.0, out_dtype=np.float32),
                                    5, 0.01, out_dtype=np.float32)
    assert out32.dtype == np.float32 and dpt32.dtype == np.float32
    _dew_close(dpt32.astype(np.float64), orc.dew_point(out32.astype(np.float64), 0.01), 0.01)


def test_wind_direction_kernel(gpu):
    from oracle import oracle as orc
    from smrf_b200 import _lib
    rng = np.random.default_rng(8)
    u, v = rng.normal(size=20000), rng.normal(size=20000)
    u[:4], v[:4] = [0.0, 1.0, 0.0, -1.0], [1.0, 0.0, -1.0, 0.0]
    az = np.empty_like(u)
    _lib.check(_lib.lib().smrfb_wind_direction(_lib.dptr(u), _lib.dptr(v), u.size, _lib.dptr(az), 0))
    ref = orc.wind_direction(u, v)
    assert np.max(np.abs(az - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-13
    assert az.min() >= 0.0 and az.max() < 360.0


def test_interp_weights_and_grid_interpolate(gpu):
    """utils.interp_weights / grid_interpolate as wind_ninja.py:171,212-235 uses them: a coarse WindNinja-like mesh onto the
    model grid (partly outside the coarse mesh: NaN), one field and a block of fields."""
    from oracle import oracle as orc
    from scipy.interpolate import LinearNDInterpolator
    from scipy.spatial import Delaunay
    from smrf_b200 import utils
    wx, wy = np.meshgrid(np.linspace(0.0, 6000.0, 31), np.linspace(500.0, 4500.0, 21))
    X, Y = np.meshgrid(np.linspace(-100.0, 6100.0, 140), np.linspace(4600.0, 400.0, 97))
    xy = np.column_stack([wx.ravel(), wy.ravel()])
    uv = np.column_stack([X.ravel(), Y.ravel()])
    vtx, wts = utils.interp_weights(xy, uv, d=2)
    ovtx, owts = orc.interp_weights(xy, uv, d=2)
    np.testing.assert_array_equal(vtx, ovtx)
    np.testing.assert_array_equal(wts, owts)
    rng = np.random.default_rng(2)
    vals = rng.normal(5.0, 2.0, size=(6, xy.shape[0]))
    one = utils.grid_interpolate(vals[0], vtx, wts, X.shape)
    ref = orc.grid_interpolate(vals[0], vtx, wts, X.shape)
    assert parity(one, ref, 1.0) <= 1e-13 and np.isnan(one).any() and not np.isnan(one).all()
    blk = utils.grid_interpolate(vals, vtx, wts, X.shape, fill_value=-9.0)
    for t in range(6):
        assert parity(blk[t], orc.grid_interpolate(vals[t], vtx, wts, X.shape, fill_value=-9.0), 1.0) <= 1e-13
    tri = Delaunay(xy)
    got = utils.grid_interpolate_deconstructed(tri, vals[1], (X, Y), method="linear")
    want = LinearNDInterpolator(tri, vals[1])((X, Y))
    assert parity(got, want, 1.0) <= 1e-12
    with pytest.raises(NotImplementedError):
        utils.grid_interpolate_deconstructed(tri, vals[1], (X, Y), method="cubic")


class _Topo:
    pass


class _Data:
    pass


def _basin(S=12, T=8, seed=3):
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(48, 64, dx=30.0, seed=seed)
    topo = _Topo()
    topo.X, topo.Y = np.meshgrid(x, y)
    topo.dem = dem
    topo.mask = np.ones_like(dem)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=seed + 1)
    names = ["ST%02d" % i for i in range(S)]
    data = _Data()
    data.metadata = pd.DataFrame({"utm_x": mx, "utm_y": my, "elevation": mz}, index=names)
    times = pd.date_range("1998-01-14 15:00", periods=T, freq="h")

    def frame(kind, sd, p=0.08):
        return pd.DataFrame(syn.dropouts_iid(syn.make_series(kind, mz, T, seed=sd), p=p, seed=sd + 50), index=times, columns=names)
    rng = np.random.default_rng(seed)
    frames = {"air_temp": frame("air_temp", 3), "vapor_pressure": frame("vapor_pressure", 4),
              "precip": np.abs(frame("air_temp", 5, p=0.0)) * 0.3,
              "wind_speed": pd.DataFrame(np.abs(rng.normal(4.0, 3.0, (T, S))), index=times, columns=names),
              "wind_direction": pd.DataFrame(rng.uniform(0.0, 360.0, (T, S)), index=times, columns=names),
              "cloud_factor": pd.DataFrame(rng.uniform(-0.1, 1.1, (T, S)), index=times, columns=names)}
    frames["precip"].iloc[1] = 0.0                                   # an hour without precipitation: skipped (precipitation.py:335)
    return topo, data, frames, times, (mx, my, mz)


CFG = {"air_temp": {"distribution": "idw", "detrend": True, "detrend_slope": -1, "idw_power": 2.0, "min": -73.0, "max": 47.0, "stations": None},
       "vapor_pressure": {"distribution": "idw", "detrend": True, "detrend_slope": -1, "idw_power": 2.0, "min": 20.0, "max": 5000.0,
                          "stations": None, "dew_point_tolerance": 0.01, "dew_point_nthreads": 2},
       "wind": {"distribution": "idw", "detrend": False, "detrend_slope": 0, "idw_power": 2.0, "min": 0.447, "max": 35.0, "stations": None},
       "precip": {"distribution": "dk", "detrend": True, "detrend_slope": 1, "dk_ncores": 2, "min": 0.0, "max": None, "stations": None},
       "cloud_factor": {"distribution": "idw", "detrend": False, "detrend_slope": 0, "idw_power": 2.0, "min": 0.0, "max": 1.0, "stations": None}}


def _oracles(topo, sta):
    from oracle import oracle as orc
    mx, my, mz = sta
    idw = orc.IDW(mx, my, topo.X, topo.Y, mz=mz, GridZ=topo.dem, power=2.0)
    dk = orc.DK(mx, my, mz, topo.X, topo.Y, topo.dem, {"dk_ncores": 2, "detrend_slope": 1})
    return orc, idw, dk


def _make_variables(topo, data, date_time):
    from smrf_b200.distribute.air_temp import ta
    from smrf_b200.distribute.cloud_factor import cf
    from smrf_b200.distribute.precipitation import ppt
    from smrf_b200.distribute.vapor_pressure import vp
    from smrf_b200.distribute.wind import Wind
    v = {"air_temp": ta(dict(CFG["air_temp"])), "vapor_pressure": vp(dict(CFG["vapor_pressure"])), "wind": Wind(dict(CFG["wind"])),
         "precip": ppt(dict(CFG["precip"])), "cloud_factor": cf(dict(CFG["cloud_factor"]))}
    for obj in v.values():
        obj.initialize(topo, data, date_time)
    return v


def test_variable_wrappers_per_step_like_the_reference_loop(gpu):
    """air_temp.py:73-86, vapor_pressure.py:104-164, wind.py:161-190 (three _distribute calls, clamp on the speed only),
    precipitation.py:335-369 (zero hours skipped, storm total through the same interpolator), cloud_factor.py:72-86."""
    topo, data, frames, times, sta = _basin()
    orc, oidw, odk = _oracles(topo, sta)
    v = _make_variables(topo, data, times)
    mz = sta[2]
    storm = frames["precip"].iloc[2:5].sum()
    v["precip"].storms = pd.DataFrame([dict(storm.to_dict(), start=times[2], end=times[4])])
    for k, t in enumerate(times):
        v["air_temp"].distribute(frames["air_temp"].loc[t])
        row = frames["air_temp"].loc[t].values
        ta_ref = orc.set_min_max(oidw.detrendedIDW(row.copy(), -1), -73.0, 47.0)
        assert parity(v["air_temp"].air_temp, ta_ref, step_scale(row)) <= TOL
        v["vapor_pressure"].distribute(frames["vapor_pressure"].loc[t], v["air_temp"].air_temp)
        row = frames["vapor_pressure"].loc[t].values
        vp_ref = orc.set_min_max(oidw.detrendedIDW(row.copy(), -1), 20.0, 5000.0)
        assert parity(v["vapor_pressure"].vapor_pressure, vp_ref, step_scale(row)) <= TOL
        dpt_ref = orc.dew_point_adjust(orc.dew_point(v["vapor_pressure"].vapor_pressure, 0.01), v["air_temp"].air_temp)
        _dew_close(v["vapor_pressure"].dew_point, dpt_ref, 0.01)
        assert (v["vapor_pressure"].dew_point < v["air_temp"].air_temp).all() and v["vapor_pressure"].precip_temp is v["vapor_pressure"].dew_point
        v["wind"].distribute(frames["wind_speed"].loc[t], frames["wind_direction"].loc[t])
        ws = frames["wind_speed"].loc[t].values
        assert parity(v["wind"].wind_speed, orc.set_min_max(oidw.calculateIDW(ws.copy()), 0.447, 35.0), step_scale(ws)) <= TOL
        wd = frames["wind_direction"].loc[t].values
        u_ref, v_ref = oidw.calculateIDW(np.sin(wd * np.pi / 180)), oidw.calculateIDW(np.cos(wd * np.pi / 180))
        assert parity(v["wind"].u_direction_distributed, u_ref, 1.0) <= TOL          # NOT clamped to [0.447, 35]
        az_ref = orc.wind_direction(u_ref, v_ref)
        d = np.abs(v["wind"].wind_direction - az_ref)
        assert np.max(np.minimum(d, 360.0 - d)) <= 1e-6                             # degrees; atan2 amplifies 1e-9 near u = v = 0
        v["cloud_factor"].distribute(frames["cloud_factor"].loc[t])
        cfr = frames["cloud_factor"].loc[t].values
        assert parity(v["cloud_factor"].cloud_factor, orc.set_min_max(oidw.calculateIDW(cfr.copy()), 0.0, 1.0), 1.0) <= TOL
        v["precip"].nasde_model = ("susong1999", "marks2017", "piecewise_susong1999")[k % 3]
        v["precip"].distribute(frames["precip"].loc[t], t, precip_temp=v["vapor_pressure"].precip_temp)
        pr = frames["precip"].loc[t].values
        if pr.sum() > 0:
            assert parity(v["precip"].precip, orc.set_min_max(odk.calculate(pr.copy()), 0.0, None), step_scale(pr)) <= TOL
            # phase / density of the new snow from the precipitation temperature (precipitation.py:371-384, 428-431)
            if v["precip"].nasde_model == "marks2017" and not (v["precip"].storming and v["vapor_pressure"].precip_temp.min() < 2.0):
                assert not v["precip"].snow_density.any() and not v["precip"].percent_snow.any()
            else:
                rho_ref, pcs_ref = orc.snow_phase_and_density(v["vapor_pressure"].precip_temp, v["precip"].precip, v["precip"].nasde_model)
                np.testing.assert_array_equal(v["precip"].percent_snow, pcs_ref)
                _snow_close(v["precip"].snow_density, rho_ref)
        else:
            assert not v["precip"].precip.any() and not v["precip"].percent_snow.any() and not v["precip"].snow_density.any()
        if k == 2:                                                                  # first hour of the storm
            st = storm.values.astype(float)
            assert parity(v["precip"].storm_total, orc.set_min_max(odk.calculate(st.copy()), 0.0, None), step_scale(st)) <= TOL


def _snow_close(got, ref):
    """pcs is table / linear arithmetic in the reference's operation order: bit for bit.  The densities go through pow / exp,
    where the device's routines differ from glibc's by <= 2 ulp: 1e-12 relative."""
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1.0), initial=0.0) <= 1e-12


def test_snow_phase_and_density_golden_and_oracle(gpu, golden):
    """smrfb_snow_phase (Snow.phase_and_density, envphys/snow.py:37-75) against the reference-generated vectors and the oracle."""
    import torch
    from oracle import oracle as orc
    from smrf_b200 import _lib
    from smrf_b200.snow import Snow
    g = golden("snow")
    for model in ("susong1999", "piecewise_susong1999", "marks2017"):
        for tag in ("pp", "storm", "zero"):
            t0 = g["t"].copy()
            rho, pcs = Snow.phase_and_density(t0, g[tag], model)
            np.testing.assert_array_equal(t0, g["t"])
            np.testing.assert_array_equal(pcs, g["%s_%s_pcs" % (model, tag)])
            _snow_close(rho, g["%s_%s_rho" % (model, tag)])
            if model == "susong1999":
                np.testing.assert_array_equal(rho, g["%s_%s_rho" % (model, tag)])
    rng = np.random.default_rng(12)
    t = rng.uniform(-14.0, 3.0, (301, 257))
    pp = np.where(rng.random(t.shape) < 0.4, 0.0, rng.gamma(0.8, 3.0, t.shape))
    t[5, 7] = np.nan
    pp[9, 9] = np.nan
    for model in ("susong1999", "piecewise_susong1999", "marks2017"):
        rho, pcs = Snow.phase_and_density(t, pp, model)
        rho_ref, pcs_ref = orc.snow_phase_and_density(t, pp, model)
        np.testing.assert_array_equal(pcs, pcs_ref)
        _snow_close(rho, rho_ref)
        # device-resident form (what a fused caller uses): same kernels
        d_t, d_pp = torch.from_numpy(t).cuda(), torch.from_numpy(pp).cuda()
        d_pcs, d_rho = torch.empty_like(d_t), torch.empty_like(d_t)
        d_any = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().smrfb_snow_phase_dev(d_t.data_ptr(), d_pp.data_ptr(), t.size, Snow.MODELS[model], d_pcs.data_ptr(),
                                                   d_rho.data_ptr(), d_any.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_pcs.cpu().numpy(), pcs)
        np.testing.assert_array_equal(d_rho.cpu().numpy(), rho)
    with pytest.raises(ValueError):
        Snow.phase_and_density(t, pp, "nope")


def test_threaded_run_loop_with_block_prefetch_and_sink(gpu, tmp_path):
    """model_framework.py:568-632: one thread per variable over date-indexed queues; with block_steps the per-step calls
    are served from fused block launches; the float32 sink receives every grid."""
    from scipy.io import netcdf_file
    from smrf_b200.output import BlockSink
    from smrf_b200.run_loop import run_threaded
    topo, data, frames, times, sta = _basin(T=6)
    x, y = topo.X[0], topo.Y[:, 0]
    res = {}
    for mode in (None, 3):
        v = _make_variables(topo, data, times)
        del v["precip"]
        sink_dir = os.path.join(str(tmp_path), "blk%s" % mode)
        names = ["air_temp", "vapor_pressure", "dew_point", "wind_speed", "wind_direction", "cloud_factor"]
        sink = BlockSink(sink_dir, names, x, y, hours=np.arange(len(times)), start_date=str(times[0]), max_pending=4)
        index = {t: i for i, t in enumerate(times)}
        q = run_threaded(v, frames, times, outputs=names, block_steps=mode,
                         sink=lambda n, t, grid: sink(n, (index[t], index[t] + 1), (0, grid.shape[0]), grid[None].copy()))
        sink.close()
        res[mode] = {n: np.stack([q[n].get(t) for t in times]) for n in names}
        f = netcdf_file(os.path.join(sink_dir, "dew_point.nc"), mmap=False)
        np.testing.assert_array_equal(f.variables["dew_point"][:], res[mode]["dew_point"].astype(np.float32))
    for n in res[None]:
        if n in ("dew_point", "wind_direction"):
            assert np.max(np.abs(res[None][n] - res[3][n])) <= 1e-5
        else:
            np.testing.assert_array_equal(res[None][n], res[3][n])   # same kernels, same arithmetic per step
