"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in the build container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only) and oracle/_ref/libkrige_ref.so (oracle/build.py build_ref()).
The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.

What runs, unmodified, from /root/reference:
  * smrf/spatial/idw.py                      (imported by path)
  * smrf/spatial/dk/dk.py                    (imported under a stub package whose
                                              ``detrended_kriging.call_grid`` forwards, with the same
                                              signature, to the reference's krige_grid() C compiled in place)
  * smrf/spatial/grid.py                     (source read into memory with its two import lines
                                              patched -- scipy.interpolate.interpnd is gone in SciPy 1.18,
                                              smrf.utils.utils needs packages absent here -- then exec'd;
                                              nothing is written to disk)
  * smrf/utils/utils.py:set_min_max          (function source extracted in memory and exec'd)
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SMRF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import build as obuild  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from smrf_b200 import synthetic as syn  # noqa: E402


def load_ref_idw():
    spec = importlib.util.spec_from_file_location("ref_idw", os.path.join(REF, "smrf/spatial/idw.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.IDW


def load_ref_dk():
    assert obuild.build_ref() is not None
    pkg = types.ModuleType("ref_dkpkg")
    pkg.__path__ = []
    shim = types.ModuleType("ref_dkpkg.detrended_kriging")

    def call_grid(ad, dgrid, elevations, weights, nthreads=1):
        return orc.call_grid(ad, dgrid, elevations, weights, nthreads, impl="ref")

    shim.call_grid = call_grid
    pkg.detrended_kriging = shim
    sys.modules["ref_dkpkg"] = pkg
    sys.modules["ref_dkpkg.detrended_kriging"] = shim
    spec = importlib.util.spec_from_file_location("ref_dkpkg.dk", os.path.join(REF, "smrf/spatial/dk/dk.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_dkpkg.dk"] = m
    spec.loader.exec_module(m)
    return m.DK


def load_ref_grid():
    src = open(os.path.join(REF, "smrf/spatial/grid.py")).read()
    src = src.replace("from scipy.interpolate.interpnd import _ndim_coords_from_arrays",
                      "from scipy.interpolate._interpnd import _ndim_coords_from_arrays")
    src = src.replace("from smrf.utils.utils import grid_interpolate_deconstructed",
                      "grid_interpolate_deconstructed = None  # grid_local path not exercised")
    m = types.ModuleType("ref_grid")
    exec(compile(src, "ref_grid.py", "exec"), m.__dict__)
    return m.GRID


def load_ref_set_min_max():
    src = open(os.path.join(REF, "smrf/utils/utils.py")).read()
    mt = re.search(r"^def set_min_max\(.*?(?=^def )", src, re.S | re.M)
    ns = {"np": np}
    exec(mt.group(0), ns)
    return ns["set_min_max"]


def rme_stations():
    md = pd.read_csv(os.path.join(REF, "smrf/tests/basins/RME/station_data/metadata.csv"), index_col=0)
    return md


def rme_series(name):
    return pd.read_csv(os.path.join(REF, "smrf/tests/basins/RME/station_data/%s.csv" % name),
                       index_col=0, parse_dates=True)


def main():
    IDW = load_ref_idw()
    DK = load_ref_dk()
    GRID = load_ref_grid()
    set_min_max = load_ref_set_min_max()

    # ------------------------------------------------------------------ IDW, RME real stations
    md = rme_stations()
    mx, my, mz = md.utm_x.values, md.utm_y.values, md.elevation.values.astype(float)
    # the real topo.nc is HDF5 and unreadable here: 10 m synthetic DEM spanning the two stations
    x, y, dem = syn.make_dem(36, 48, dx=10.0, x0=519560.0, y0=4768400.0, seed=1)
    dem = (dem - 1800.0) * 0.1 + 2070.0
    dem = dem.astype(np.float32).astype(np.float64)
    X, Y = np.meshgrid(x, y)
    ta = rme_series("air_temp")[md.index].values[:12]
    vp = rme_series("vapor_pressure")[md.index].values[:12]
    cf = rme_series("cloud_factor")[md.index].values[:12]
    ref = IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0)
    out_ta = np.stack([set_min_max(ref.detrendedIDW(r.copy(), -1), -73.0, 47.0) for r in ta])
    pv_ta = []
    for r in ta:
        ref.detrendData(r.copy(), -1)
        pv_ta.append(ref.pv.astype(float))
    out_vp = np.stack([set_min_max(ref.detrendedIDW(r.copy(), -1), 20.0, 5000.0) for r in vp])
    out_cf = np.stack([set_min_max(ref.calculateIDW(r.copy()), 0.0, 1.0) for r in cf])
    np.savez_compressed(os.path.join(OUT, "idw_rme.npz"), mx=mx, my=my, mz=mz, x=x, y=y, dem=dem,
                        air_temp=ta, vapor_pressure=vp, cloud_factor=cf,
                        out_air_temp=out_ta, out_vapor_pressure=out_vp, out_cloud_factor=out_cf,
                        pv_air_temp=np.array(pv_ta))

    # ------------------------------------------------------------------ IDW, synthetic with dropouts
    x, y, dem = syn.make_dem(24, 31, dx=30.0, seed=2)
    X, Y = np.meshgrid(x, y)
    mx, my, mz = syn.make_stations(20, x, y, dem, seed=7)
    data = syn.dropouts_iid(syn.make_series("air_temp", mz, 10, seed=3), p=0.15, seed=11)
    data[4, 1:] = np.nan                    # a single valid station (polyfit min-norm case)
    data[5, :] = 3.25                       # constant data
    data[6, :] = 0.0                        # all zeros
    res = {}
    for power in (2.0, 1.5):
        ref = IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=power)
        for flag in (-1, 0, 1):
            outs, pvs = [], []
            for r in data:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    outs.append(ref.detrendedIDW(r.copy(), flag))
                pvs.append(np.asarray(ref.pv, dtype=float))
            res["out_p%g_f%d" % (power, flag)] = np.stack(outs)
            res["pv_p%g_f%d" % (power, flag)] = np.stack(pvs)
        res["out_p%g_plain" % power] = np.stack([ref.calculateIDW(r.copy()) for r in data])
    # a station exactly on a cell centre (zero distance => inf weight => NaN cell)
    mx0, my0 = mx.copy(), my.copy()
    mx0[3], my0[3] = x[5], y[7]
    ref = IDW(mx0, my0, X, Y, mz=mz, GridZ=dem, power=2.0)
    with np.errstate(all="ignore"):
        res["out_zero_dist"] = np.stack([ref.detrendedIDW(r.copy(), 0) for r in data[:4]])
    np.savez_compressed(os.path.join(OUT, "idw_synth.npz"), mx=mx, my=my, mz=mz, x=x, y=y, dem=dem, data=data,
                        mx0=mx0, my0=my0, **res)

    # ------------------------------------------------------------------ DK, synthetic
    x, y, dem = syn.make_dem(12, 15, dx=90.0, seed=4)
    X, Y = np.meshgrid(x, y)
    mx, my, mz = syn.make_stations(12, x, y, dem, seed=8)
    pr = syn.make_series("precip", mz, 40, seed=5)
    pr = pr[pr.sum(axis=1) > 0][:6]
    pr[2, 3] = np.nan
    pr[3, 3] = np.nan
    pr[4, [0, 7]] = np.nan
    cfg = {"dk_ncores": 2, "detrend_slope": 1}
    ref = DK(mx, my, mz, X, Y, dem, cfg)
    outs, wts, pvs, masks = [], [], [], []
    for r in pr:
        v = ref.calculate(r.copy())
        outs.append(set_min_max(v, 0.0, None))
        full = np.zeros(X.shape + (len(mx),))
        full[:, :, ~ref.nan_val] = ref.weights
        wts.append(full)
        pvs.append(np.asarray(ref.pv, dtype=float))
        masks.append(ref.nan_val.copy())
    np.savez_compressed(os.path.join(OUT, "dk_synth.npz"), mx=mx, my=my, mz=mz, x=x, y=y, dem=dem, data=pr,
                        out=np.stack(outs), weights=np.stack(wts), pv=np.stack(pvs), nan_mask=np.stack(masks))

    # ------------------------------------------------------------------ DK weights KAT: reference C direct
    rng = np.random.default_rng(21)
    n, ng = 24, 150
    sx, sy = rng.uniform(0, 20000, n), rng.uniform(0, 20000, n)
    sz = rng.uniform(1200, 2600, n)
    cx, cy = rng.uniform(0, 20000, ng), rng.uniform(0, 20000, ng)
    ad = orc.station_distances(sx, sy)
    dg = np.sqrt((cx[:, None] - sx[None, :]) ** 2 + (cy[:, None] - sy[None, :]) ** 2)
    w = np.zeros_like(dg)
    orc.call_grid(ad, dg, sz, w, 2, impl="ref")
    np.savez_compressed(os.path.join(OUT, "krige_kat.npz"), sx=sx, sy=sy, sz=sz, cx=cx, cy=cy, ad=ad, dgrid=dg, weights=w)

    # ------------------------------------------------------------------ DK, RME precip (2 stations)
    md = rme_stations()
    mx, my, mz = md.utm_x.values, md.utm_y.values, md.elevation.values.astype(float)
    x, y, dem = syn.make_dem(20, 26, dx=20.0, x0=519500.0, y0=4768450.0, seed=6)
    dem = ((dem - 1800.0) * 0.1 + 2070.0).astype(np.float32).astype(np.float64)
    X, Y = np.meshgrid(x, y)
    pr = rme_series("precip")[md.index].values
    pr = pr[pr.sum(axis=1) > 0][:8]
    ref = DK(mx, my, mz, X, Y, dem, {"dk_ncores": 2, "detrend_slope": 1})
    out = np.stack([set_min_max(ref.calculate(r.copy()), 0.0, None) for r in pr])
    np.savez_compressed(os.path.join(OUT, "dk_rme.npz"), mx=mx, my=my, mz=mz, x=x, y=y, dem=dem, data=pr, out=out)

    # ------------------------------------------------------------------ GRID linear
    x, y, dem = syn.make_dem(40, 50, dx=100.0, seed=9)
    X, Y = np.meshgrid(x, y)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=900.0, jitter=60.0, seed=5)
    # drop a corner of the lattice so part of the DEM lies outside the convex hull (NaN cells)
    keep = ~((px > x[-1] - 1500.0) & (py > y[0] - 1200.0))
    px, py, pz = px[keep], py[keep], pz[keep]
    fields = syn.make_source_fields(px, py, pz, 5, seed=9)
    mask = ((X - x.mean()) ** 2 + (Y - y.mean()) ** 2) < 1500.0 ** 2
    res = {}
    for gm in (True, False):
        cfg = {"grid_local": False, "grid_local_n": 25, "grid_mask": gm}
        ref = GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, mask=mask)
        res["ptmask_%d" % gm] = ref.mask.copy()
        for flag in (-1, 0):
            outs, pvs = [], []
            for r in fields:
                outs.append(ref.detrendedInterpolation(pd.Series(r.copy()), flag, "linear"))
                pvs.append(np.asarray(ref.pv, dtype=float))
            res["out_m%d_f%d" % (gm, flag)] = np.stack(outs)
            res["pv_m%d_f%d" % (gm, flag)] = np.stack(pvs)
    res["out_plain"] = np.stack([ref.calculateInterpolation(r.copy(), "linear") for r in fields])
    res["out_plain_nearest"] = np.stack([ref.calculateInterpolation(r.copy(), "nearest") for r in fields])
    res["out_m0_f-1_nearest"] = np.stack([ref.detrendedInterpolation(pd.Series(r.copy()), -1, "nearest") for r in fields])
    np.savez_compressed(os.path.join(OUT, "grid_synth.npz"), px=px, py=py, pz=pz, x=x, y=y, dem=dem, mask=mask,
                        fields=fields, **res)

    # ------------------------------------------------------------------ clamp
    rng = np.random.default_rng(1)
    a = rng.normal(0, 3, (7, 9))
    a[2, 3] = np.nan
    a[0, 0] = np.inf
    np.savez_compressed(os.path.join(OUT, "clamp.npz"), a=a,
                        out_both=set_min_max(a.copy(), -1.0, 2.0),
                        out_min=set_min_max(a.copy(), 0.0, None),
                        out_none=set_min_max(a.copy(), None, None))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
