#!/usr/bin/env python
"""bench.py -- distributed cell-timesteps/sec of the SMRF spatial-distribution hot path on B200.

Workloads (BASELINE.json configs):
  c4 (default, configs[3]): synthetic 10 000 x 10 000 DEM (1e8 cells, 30 m), 200 stations, hourly series of one year;
      one *step* = one block of T_b = 720 hourly timesteps (30 days) distributed for three fields over the whole grid,
      walked slab by slab (1024 DEM rows per slab) so that the device-resident output block [T_b][slab] stays at 59 GB:
          air_temp        IDW, detrended (slope -1), clamp [-73, 47], i.i.d. 2 % station dropouts
          vapor_pressure  IDW, detrended (slope -1), clamp [20, 5000], i.i.d. 2 % station dropouts
          precip          DK,  detrended (slope +1), clamp [0, inf), block outages (<= 64 masks / year)
  c5 (configs[4]): synthetic 4 000 x 4 000 DEM (10 m) fed by a 71 x 71 lattice of 3 km source points (HRRR-like) + IDW:
          air_temp        GRID linear, detrended by the elevation trend of the points (slope -1), clamp [-73, 47]
          wind_speed      GRID linear, no detrend, clamp [0.447, 35]
          vapor_pressure  GRID grid_local (per-point local lapse rates, three interpolations fused), clamp [20, 5000]
          precip          IDW over 200 stations scaled to the domain, detrended (+1), clamp [0, inf)
value = cells * T_b * fields * steps / time with the inputs resident in HBM and the output written into a device-resident
block (CUDA events on the launch stream, max over ranks).  One-off work per geometry / NaN mask (DK kriging weights, GRID
triangulation, local fits of grid_local) is done during warm-up and reported separately; value_year_amortised folds the
DK weight builds of a whole year back in.
e2e = the same fields through the product-level driver (smrf_b200.sharding.BlockRunner over the drop-in classes) with
HOST buffers: station rows H2D and the distributed grids D2H into pinned memory inside the timed region.
Multi-GPU (torchrun, one rank per GPU): timestep blocks shard across ranks, no collective on the data path; weak scaling.

--impl reference times the reference's CPU path (NumPy IDW as idw.py does it, scipy griddata as grid.py calls it, the
reference's krige.c compiled in place when oracle/_ref is present) on a bounded tile of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INF = float("inf")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"])
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--stations", type=int, default=200)
    ap.add_argument("--tb", type=int, default=0, help="timesteps per block (c4: 720 = 30 days of hourly steps; c5: 512)")
    ap.add_argument("--slab-rows", type=int, default=1024, help="DEM rows per cell slab (c4); the device-resident output block is [tb][slab]")
    ap.add_argument("--te2e", type=int, default=0, help="timesteps per block on the host-buffer (e2e) leg; 0 = max(16, 32 // n_gpus)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the float32 and per-step-call measurements")
    a = ap.parse_args()
    if a.workload == "c4":
        a.ny, a.nx, a.tb = a.ny or 10000, a.nx or 10000, a.tb or 720
    else:
        a.ny, a.nx, a.tb = a.ny or 4000, a.nx or 4000, a.tb or 512
    return a


# ------------------------------------------------------------------------------------------------
# synthetic basins
# ------------------------------------------------------------------------------------------------
def make_dem(ny, nx, dx):
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(ny, nx, dx=dx, noise=(ny * nx <= 25_000_000))
    if ny * nx > 25_000_000:      # cheap deterministic micro-relief instead of 1e8 normal draws
        j = np.arange(nx, dtype=np.float64)[None, :]
        i = np.arange(ny, dtype=np.float64)[:, None]
        dem = (dem + 8.0 * np.sin(0.7 * j + 1.3 * i) * np.cos(0.31 * i - 0.2 * j)).astype(np.float32).astype(np.float64)
    return x, y, dem


def make_c4(ny, nx, S, T_total):
    from smrf_b200 import synthetic as syn
    x, y, dem = make_dem(ny, nx, 30.0)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
    series = {"air_temp": syn.dropouts_iid(syn.make_series("air_temp", mz, T_total, seed=3), p=0.02, seed=11),
              "vapor_pressure": syn.dropouts_iid(syn.make_series("vapor_pressure", mz, T_total, seed=4), p=0.02, seed=12)}
    pr, nmask = syn.dropouts_block(syn.make_series("precip", mz, T_total, seed=5), max_masks=12, seed=13)   # <= 64 allowed
    series["precip"] = pr
    return x, y, dem, (mx, my, mz), series, nmask


def make_c5(ny, nx, S, T_total):
    from smrf_b200 import synthetic as syn
    x, y, dem = make_dem(ny, nx, 10.0)
    rng = np.random.default_rng(5)
    cx, cy = 0.5 * (x[0] + x[-1]), 0.5 * (y[0] + y[-1])
    g = 3000.0 * (np.arange(71) - 35.0)                       # 71 x 71 points on a 3 km lattice centred on the DEM
    PX, PY = np.meshgrid(cx + g, cy + g)
    px = PX.ravel() + rng.uniform(-150.0, 150.0, PX.size)
    py = PY.ravel() + rng.uniform(-150.0, 150.0, PX.size)
    j = np.clip(np.rint((px - x[0]) / (x[1] - x[0])).astype(int), 0, nx - 1)
    i = np.clip(np.rint((py - y[0]) / (y[1] - y[0])).astype(int), 0, ny - 1)
    pz = 1800.0 + 700.0 * np.sin(2 * np.pi * px / 41000.0) * np.cos(2 * np.pi * py / 57000.0) + rng.normal(0.0, 30.0, px.size)
    inside = (px >= x[0]) & (px <= x[-1]) & (py <= y[0]) & (py >= y[-1])
    pz[inside] = dem[i[inside], j[inside]] + rng.normal(0.0, 30.0, int(inside.sum()))
    f = syn.make_source_fields(px, py, pz, T_total, seed=9)
    series = {"air_temp": f,
              "wind_speed": 0.5 + 0.35 * np.abs(syn.make_source_fields(px, py, pz, T_total, seed=10)),
              "vapor_pressure": 600.0 + 40.0 * syn.make_source_fields(px, py, pz, T_total, seed=11)}
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
    series["precip"] = syn.dropouts_iid(np.abs(syn.make_series("air_temp", mz, T_total, seed=6)), p=0.02, seed=14)
    return x, y, dem, (px, py, pz), (mx, my, mz), series


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on a bounded tile of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_arm(workload, S, steps, warmup, t_per_step=2):
    from oracle import oracle as orc
    from oracle import build as obuild
    from smrf_b200 import synthetic as syn
    obuild.build_oracle()
    ncores = os.cpu_count() or 1
    T = (steps + warmup) * t_per_step
    if workload == "c4":
        has_ref = orc.krige_ref_lib() is not None
        impl = "ref" if has_ref else "oracle"
        x, y, dem = syn.make_dem(2000, 2000, dx=30.0, noise=False)
        mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
        ta = syn.dropouts_iid(syn.make_series("air_temp", mz, T, seed=3), p=0.02, seed=11)
        vp = syn.dropouts_iid(syn.make_series("vapor_pressure", mz, T, seed=4), p=0.02, seed=12)
        pr = syn.make_series("precip", mz, 400, seed=5)
        pr = pr[pr.sum(axis=1) > 0][:T]
        while len(pr) < T:
            pr = np.concatenate([pr, pr])[:T]
        iy = ix = 300
        Xi, Yi = np.meshgrid(x[800:800 + ix], y[700:700 + iy])
        Zi = dem[700:700 + iy, 800:800 + ix]
        ky = kx = 20
        Xk, Yk = np.meshgrid(x[800:800 + kx], y[700:700 + ky])
        Zk = dem[700:700 + ky, 800:800 + kx]
        t0 = time.perf_counter()
        idw = orc.IDW(mx, my, Xi, Yi, mz=mz, GridZ=Zi, power=2.0)          # the [ny,nx,S] cubes, as idw.py:53-77
        t_idw_setup = time.perf_counter() - t0
        dk = orc.DK(mx, my, mz, Xk, Yk, Zk, {"dk_ncores": ncores, "detrend_slope": 1}, impl=impl)
        t0 = time.perf_counter()
        dk.calculate(pr[0].copy())                                          # builds the kriging weights (one mask)
        t_dk_build = time.perf_counter() - t0
        # steady state of DK on the SAME 300 x 300 tile as IDW: the 20 x 20 tile's weights repeated (weight values do not
        # change the cost of dk.py:91-95's weighted sum; building 90 000 cells' weights on the CPU would take minutes)
        dkb = orc.DK(mx, my, mz, Xi, Yi, Zi, {"dk_ncores": ncores, "detrend_slope": 1}, impl=impl)
        dkb.weights = np.tile(dk.weights, (iy // ky, ix // kx, 1))
        dkb.nan_val = dk.nan_val.copy()

        def f_idw(rows, lo, hi):
            for r in rows:
                orc.set_min_max(idw.detrendedIDW(r.copy(), -1), lo, hi)

        def f_dk(rows):
            for r in rows:                                      # no dropouts in this sample: the cached mask always applies
                orc.set_min_max(dkb.calculate(r.copy()), 0.0, None)

        def threads(sl):
            return [threading.Thread(target=f_idw, args=(ta[sl], -73.0, 47.0)),
                    threading.Thread(target=f_idw, args=(vp[sl], 20.0, 5000.0)),
                    threading.Thread(target=f_dk, args=(pr[sl],))]
        cells_per_hour = 3 * iy * ix
        kind = "port"
        sample = ("%d steps x %d hours: IDW air_temp + vapor_pressure and DK precip (steady state) on one %dx%d tile, %d stations, "
                  "3 Python threads (one per variable, as model_framework.py:615-632), NumPy cubes as idw.py / dk.py build them "
                  "(NumPy restatement of the reference's Python; kriging weights from %s on a %dx%d tile with %d OpenMP "
                  "threads, repeated over the tile): weight build %.1f s = %.1f cells/s, excluded like on the GPU; IDW cube setup %.1f s"
                  % (steps, t_per_step, iy, ix, S, "the reference's krige.c compiled in place" if has_ref else "the C restatement of krige.c",
                     ky, kx, ncores, t_dk_build, ky * kx / t_dk_build, t_idw_setup))
        extra = {"dk_weight_build_cells_per_s": ky * kx / t_dk_build}
    else:
        x, y, dem, (px, py, pz), (mx, my, mz), series = make_c5(4000, 4000, S, T)
        iy = ix = 300
        r0, c0 = 1800, 1900
        Xi, Yi = np.meshgrid(x[c0:c0 + ix], y[r0:r0 + iy])
        Zi = dem[r0:r0 + iy, c0:c0 + ix]
        import pandas as pd
        md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                           "longitude": -116.0 + (px - px.min()) / 82000.0})
        g_lin = orc.GRID({"grid_local": False, "grid_mask": False}, px, py, Xi, Yi, mz=pz, GridZ=Zi)
        g_loc = orc.GRID({"grid_local": True, "grid_local_n": 25, "grid_mask": False}, px, py, Xi, Yi, mz=pz, GridZ=Zi, metadata=md)
        idw = orc.IDW(mx, my, Xi, Yi, mz=mz, GridZ=Zi, power=2.0)

        def f_lin(rows, det, lo, hi):
            for r in rows:
                v = g_lin.detrendedInterpolation(r.copy(), -1, "linear") if det else g_lin.calculateInterpolation(r.copy(), "linear")
                orc.set_min_max(v, lo, hi)

        def f_loc(rows):
            for r in rows:
                orc.set_min_max(g_loc.detrendedInterpolation(r.copy(), -1, "linear"), 20.0, 5000.0)

        def f_idw(rows):
            for r in rows:
                orc.set_min_max(idw.detrendedIDW(r.copy(), 1), 0.0, None)

        def threads(sl):
            return [threading.Thread(target=f_lin, args=(series["air_temp"][sl], True, -73.0, 47.0)),
                    threading.Thread(target=f_lin, args=(series["wind_speed"][sl], False, 0.447, 35.0)),
                    threading.Thread(target=f_loc, args=(series["vapor_pressure"][sl],)),
                    threading.Thread(target=f_idw, args=(series["precip"][sl],))]
        cells_per_hour = 4 * iy * ix
        kind = "port"
        sample = ("%d steps x %d hours: GRID linear (detrended, plain), GRID grid_local and IDW on one %dx%d tile of the C5 basin, %d "
                  "source points, %d stations; scipy.interpolate.griddata / LinearNDInterpolator called per step as grid.py:179-231 and "
                  ":115-177 do (the triangulation is rebuilt on every griddata call, as in the reference), 4 Python threads"
                  % (steps, t_per_step, iy, ix, len(px), S))
        extra = {}

    def one_step(k):
        th = threads(slice(k * t_per_step, (k + 1) * t_per_step))
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for k in range(warmup):
        one_step(k)
    tot = sum(one_step(warmup + k) for k in range(steps))
    cb = {"value": steps * t_per_step * cells_per_hour / tot, "unit": "cell-timesteps/s", "cores": ncores, "kind": kind, "sample": sample}
    cb.update(extra)
    return cb, tot / steps * 1e3


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm = float(json.load(open(p))["hbm_gbs"])
            src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    fp64 = 33.8          # builder-measured DFMA issue loop (profiles/r01_peak_probe.json); MEASURED_PEAKS.json holds no fp64 figure
    q = os.path.join(ROOT, "profiles", "r01_peak_probe.json")
    if os.path.exists(q):
        try:
            fp64 = float(json.load(open(q)).get("fp64_dfma_tflops", fp64))
        except Exception:
            pass
    return hbm, src, fp64


def ncu_traffic(kernel, shape_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` at exactly this shape, from the committed
    `ncu --set full` captures (profiles/r02_traffic.json); None for shapes that were not captured."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        for e in json.load(open(p))["captures"]:
            if e["kernel"] == kernel and e["shape"] == shape_key:
                return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    metric = "distributed cell-timesteps/sec (IDW+DK)" if wl == "c4" else "distributed cell-timesteps/sec (GRID+IDW)"

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, ms = cpu_arm(wl, args.stations, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": cb["value"],
                "unit": "cell-timesteps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": ("C4 synthetic basin (BASELINE.json configs[3]): IDW air_temp + vapor_pressure, DK precip, %d stations"
                                        if wl == "c4" else
                                        "C5 synthetic basin (BASELINE.json configs[4]): GRID linear x 2, GRID grid_local, IDW, %d stations")
                                       % args.stations + "; CPU tile sample of it (the reference's [ny,nx,S] cubes need 320 GB at 1e8 cells)"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from smrf_b200 import _lib
    from smrf_b200.sharding import BlockRunner, FieldSpec
    from smrf_b200.spatial import DK, GRID, IDW
    L = _lib.lib()

    ny, nx, S, Tb = args.ny, args.nx, args.stations, args.tb
    ncell = ny * nx
    T_total = 8760
    X_b = Y_b = None
    t0 = time.perf_counter()
    if wl == "c4":
        x, y, dem, sta, series, nmask_year = make_c4(ny, nx, S, T_total)
        pts = None
        fields = [FieldSpec("air_temp", "idw", True, -1, -73.0, 47.0), FieldSpec("vapor_pressure", "idw", True, -1, 20.0, 5000.0),
                  FieldSpec("precip", "dk", True, 1, 0.0, None)]
        seg = T_total // 12                  # DK outage masks change every 730 h: block r sits inside segment r
        t_lo = (rank * seg) % (T_total - Tb)
    else:
        x, y, dem, pts, sta, series = make_c5(ny, nx, S, T_total)
        import pandas as pd
        px, py, pz = pts
        md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                           "longitude": -116.0 + (px - px.min()) / 82000.0})
        fields = [FieldSpec("air_temp", "grid", True, -1, -73.0, 47.0), FieldSpec("wind_speed", "grid", False, 0, 0.447, 35.0),
                  FieldSpec("vapor_pressure", "grid", True, -1, 20.0, 5000.0, grid_local=True, grid_local_n=25, metadata=md),
                  FieldSpec("precip", "idw", True, 1, 0.0, None)]
        nmask_year = 0
        t_lo = (rank * Tb) % (T_total - Tb)
    t_inputs = time.perf_counter() - t0
    mx, my, mz = sta
    rows = {f.name: np.ascontiguousarray(series[f.name][t_lo:t_lo + Tb]) for f in fields}
    # GridX / GridY as np.meshgrid(x, y) would give them, but as zero-copy broadcast views (1.6 GB saved at 1e8 cells)
    X_b = np.broadcast_to(x[None, :], (ny, nx))
    Y_b = np.broadcast_to(y[:, None], (ny, nx))
    objs = {}
    t0 = time.perf_counter()
    for f in fields:
        if f.method == "idw":
            objs[f.name] = IDW(mx, my, X_b, Y_b, mz=mz, GridZ=dem, power=2.0, device=local_rank)
        elif f.method == "dk":
            objs[f.name] = DK(mx, my, mz, X_b, Y_b, dem, {"dk_ncores": 1, "detrend_slope": f.slope}, device=local_rank)
        else:
            cfg = {"grid_local": bool(f.opts.get("grid_local")), "grid_mask": False, "grid_local_n": f.opts.get("grid_local_n", 25)}
            objs[f.name] = GRID(cfg, pts[0], pts[1], X_b, Y_b, mz=pts[2], GridZ=dem, metadata=f.opts.get("metadata"), device=local_rank)
    t_create = time.perf_counter() - t0

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_rows = {}
    t_local_fit = 0.0
    for f in fields:
        if f.method == "grid" and f.opts.get("grid_local"):
            t0 = time.perf_counter()
            res, slope, icpt = objs[f.name].local_fit(rows[f.name], flag=f.slope, exact=False)     # host, per block (grid.py:125-141)
            t_local_fit = time.perf_counter() - t0
            d_rows[f.name] = torch.from_numpy(np.ascontiguousarray(np.stack([res, slope, icpt], axis=1))).cuda()
        else:
            d_rows[f.name] = torch.from_numpy(rows[f.name]).cuda()
    if wl == "c4":
        slab_rows = min(args.slab_rows, ny)
        slabs = [(r0, min(r0 + slab_rows, ny)) for r0 in range(0, ny, slab_rows)]
    else:
        slabs = [(0, ny)]
    slab_cells = max(b - a for a, b in slabs) * nx
    d_out = torch.empty((Tb, slab_cells), dtype=torch.float64, device="cuda")
    n_dk_masks = len({np.isnan(r).tobytes() for r in rows["precip"]}) if wl == "c4" else 0

    def launch(f, slab):
        lo = -INF if f.vmin is None else f.vmin
        hi = INF if f.vmax is None else f.vmax
        cbeg, ccnt = slab[0] * nx, (slab[1] - slab[0]) * nx
        h = objs[f.name]._h
        if f.method == "idw":
            _lib.check(L.smrfb_idw_distribute_window_dev(h, d_rows[f.name].data_ptr(), Tb, int(f.detrend), f.slope, lo, hi,
                                                         None, None, cbeg, ccnt, d_out.data_ptr(), st))
        elif f.method == "dk":
            _lib.check(L.smrfb_dk_distribute_window_dev(h, d_rows[f.name].data_ptr(), _lib.dptr(rows[f.name]), Tb, lo, hi,
                                                        None, None, cbeg, ccnt, d_out.data_ptr(), st))
        elif f.opts.get("grid_local"):
            _lib.check(L.smrfb_grid_set_local_exact(h, 0))       # block path: closed-form fits, folded + fused interpolation
            _lib.check(L.smrfb_grid_distribute_local_dev(h, d_rows[f.name].data_ptr(), Tb, lo, hi, d_out.data_ptr(), st))
        else:
            _lib.check(L.smrfb_grid_distribute_dev(h, d_rows[f.name].data_ptr(), Tb, int(f.detrend), f.slope, 0, lo, hi,
                                                   None, None, d_out.data_ptr(), st))

    ev_field = {}

    def device_step(record=None):
        for slab in slabs:
            for f in fields:
                if record is not None:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                launch(f, slab)
                if record is not None:
                    e1.record(stream)
                    record.setdefault(f.name, []).append((e0, e1, (slab[1] - slab[0]) * nx))

    # DK weights: one-off per NaN mask, built for every cell of the handle on the first call
    t_dk_build, dk_stats = 0.0, (0, 0, 0)
    if wl == "c4":
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        launch(fields[2], slabs[0])
        torch.cuda.synchronize()
        t_dk_build = time.perf_counter() - t0          # the first call builds the mask's weights for the WHOLE grid

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    if wl == "c4":
        dk_stats = objs["precip"].stats()
    idw_names = [f.name for f in fields if f.method == "idw"]
    far = all(objs[n].kernel_info()["far_field"] for n in idw_names)
    for n in idw_names:
        _lib.check(L.smrfb_idw_far_timing(objs[n]._h, 1, None))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for _ in range(args.steps):
        device_step(ev_field)
    e_end.record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = e_start.elapsed_time(e_end)
    clocks = sampler.stop() if sampler else None
    far_ms = np.zeros(3)
    for n in idw_names:
        t3 = np.zeros(3)
        _lib.check(L.smrfb_idw_far_timing(objs[n]._h, 0, _lib.dptr(t3)))
        far_ms += t3
    tms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    ms_step = ms_total / args.steps
    nf = len(fields)
    value = world * ncell * Tb * nf * args.steps / (ms_total * 1e-3)
    # per field: ms for one pass over the whole grid (all slabs), and per full-size launch (one slab)
    field_ms = {k: float(sum(a.elapsed_time(b) for a, b, _ in v)) / args.steps for k, v in ev_field.items()}
    launch_ms = {k: float(np.mean([a.elapsed_time(b) for a, b, c in v if c == slab_cells])) for k, v in ev_field.items()}

    # ---- parity sample: outputs of one more (untimed) launch per field against the CPU oracle (checker only)
    parity = None
    if rank == 0:
        try:
            parity = parity_sample(fields, objs, launch, slabs[0], d_out, rows, x, y, dem, sta, pts, nx, Tb, torch)
        except Exception as e:        # the bench line must still be printed
            parity = {"error": repr(e)[:300]}

    # ---- float32 output mode (the reference's NetCDF type, contract 1e-5): a separate figure, never the headline
    f32 = None
    if not args.no_extras and rank == 0:
        try:
            f32 = f32_mode(fields, objs, launch, slabs[0], d_out, Tb, nx, L, _lib, torch, stream)
        except Exception as e:
            f32 = {"error": repr(e)[:300]}

    # ---- e2e: the product-level driver over the public classes, host buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    del d_out
    torch.cuda.empty_cache()          # the host-buffer API keeps its own device ring per handle
    if not args.no_e2e:
        # 0.8 GB of pinned host memory per step at 1e8 cells: 25.6 GB at one GPU, 12.8 GB per rank at 2-8 (the box has 196 GB)
        Te = min(args.te2e if args.te2e > 0 else max(16, 32 // world), Tb)
        runner = BlockRunner(fields, x, y, dem, {"idw": sta, "dk": sta, "grid": pts}, rank=0, world=1, partition="time",
                             block_steps=Te, slab_rows=ny, device=local_rank, handles={f.name: objs[f.name] for f in fields},
                             nbuf=1)
        tables = {k: v[:Te] for k, v in rows.items()}
        runner.run(tables)            # warm-up: pinned buffers, rings, DK weights of every slab
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            done = runner.run(tables)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h_bytes = int(nf * Te * ncell * 8)
        ceiling = d2h_ceiling(torch) if rank == 0 else None
        ceiling_all = None
        if world > 1:
            # what the box's host side takes when every rank copies at once (the GPUs share PCIe switches and one NUMA node):
            # world x 4 GiB device -> pinned host, all ranks between two barriers, bytes / the slowest rank's time
            try:
                src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
                dst = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
                dst.copy_(src)
                barrier()
                t1 = time.perf_counter()
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                tc = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
                ceiling_all = world * 4 * float(1 << 30) / float(tc.item()) / 1e9
                del src, dst
            except Exception:
                ceiling_all = None
        e2e = {"value": world * done * n_e2e / dt, "unit": "cell-timesteps/s",
               "h2d_bytes_per_step": int(nf * Te * rows[fields[0].name].shape[1] * 8), "d2h_bytes_per_step": d2h_bytes,
               "steps_per_block": Te, "blocks_timed": n_e2e, "d2h_gbs_per_rank": d2h_bytes * n_e2e / dt / 1e9,
               "d2h_ceiling_gbs_one_rank_alone": ceiling,
               "d2h_ceiling_gbs_all_ranks_at_once": ceiling_all,
               "frac_of_d2h_ceiling": ((world * d2h_bytes * n_e2e / dt / 1e9 / ceiling_all) if ceiling_all else
                                       (d2h_bytes * n_e2e / dt / 1e9 / ceiling) if ceiling else None),
               "api": "smrf_b200.sharding.BlockRunner.run(tables) over smrf_b200.spatial.IDW/DK/GRID.distribute_block -> pinned float64 "
                      "[T, rows, nx] blocks handed to a sink, slab by slab"}
        del runner
        if not args.no_extras:
            # secondary figure: the same leg with float32 blocks (the reference's output-file type, contract 1e-5) -- half the
            # bytes over PCIe; never the headline
            try:
                r32 = BlockRunner(fields, x, y, dem, {"idw": sta, "dk": sta, "grid": pts}, rank=0, world=1, partition="time",
                                  block_steps=Te, slab_rows=ny, device=local_rank, handles={f.name: objs[f.name] for f in fields},
                                  nbuf=1, out_dtype=np.float32)
                r32.run(tables)
                barrier()
                t0 = time.perf_counter()
                for _ in range(n_e2e):
                    done32 = r32.run(tables)
                barrier()
                d32 = time.perf_counter() - t0
                t32 = torch.tensor([d32], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(t32, op=dist.ReduceOp.MAX)
                d32 = float(t32.item())
                e2e["f32_mode"] = {"value": world * done32 * n_e2e / d32, "unit": "cell-timesteps/s", "out_dtype": "f32", "contract": 1e-05,
                                   "d2h_bytes_per_step": d2h_bytes // 2, "d2h_gbs_per_rank": d2h_bytes / 2 * n_e2e / d32 / 1e9}
                del r32
            except Exception as e:
                e2e["f32_mode"] = {"error": repr(e)[:300]}

    # ---- the per-timestep drop-in call (what an unmodified run loop makes): T = 1 through image_data._distribute
    per_step = None
    if not args.no_extras and rank == 0:
        try:
            per_step = per_step_call(wl, x, y, dem, sta, pts, series, fields, local_rank)
        except Exception as e:
            per_step = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, hbm_src, fp64_peak = load_peaks()
    roof = roofline(wl, fields, objs, launch_ms, far, far_ms, args.steps, len(slabs), slab_cells, Tb, S, dk_stats, hbm_peak, hbm_src,
                    fp64_peak, ncell, pts)
    cb = None
    if not args.no_cpu_baseline and world == 1:      # reported baseline, rank 0 at N = 1 only
        cb, _ = cpu_arm(wl, S, 2, 1)
    year = None
    if wl == "c4":
        # a whole year on this rank: 8760 h of the three fields + the kriging-weight build of every NaN mask of the year
        t_dist = 8760.0 / Tb * ms_step * 1e-3
        t_build = nmask_year * t_dk_build
        year = {"value": ncell * 8760.0 * nf / (t_dist + t_build), "unit": "cell-timesteps/s", "dk_masks_per_year": nmask_year,
                "distribution_s": t_dist, "dk_weight_builds_s": t_build,
                "note": "per GPU; all %d distinct DK station masks of the synthetic year built once each (%.1f s per mask over the "
                        "whole grid), nothing else amortised" % (nmask_year, t_dk_build)}
    line = {"metric": metric, "value": value, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "arithmetic": "fp64 throughout.  IDW on uniform meshes: near stations summed directly, far stations through a 10 x 10 "
                          "Chebyshev interpolant of their weight sums per 128 x 128-cell patch (a-priori error bound 2e-10 of the step's "
                          "largest residual, observed ~1e-11; csrc/idw_far.cu); DK / GRID: exact per-cell weights",
            "config": {"workload": ("C4 (BASELINE.json configs[3]): %dx%d DEM = %d cells, %d stations, hourly; step = %d-hour block x 3 fields "
                                    "(air_temp IDW, vapor_pressure IDW, precip DK) over the whole grid, walked in %d slabs of <= %d cells"
                                    % (ny, nx, ncell, S, Tb, len(slabs), slab_cells)) if wl == "c4" else
                                   ("C5 (BASELINE.json configs[4]): %dx%d DEM = %d cells fed by %d source points on a 3 km lattice + %d stations; "
                                    "step = %d-hour block x 4 fields (air_temp GRID linear detrended, wind_speed GRID linear, vapor_pressure "
                                    "GRID grid_local, precip IDW)" % (ny, nx, ncell, len(pts[0]), S, Tb)),
                       "timesteps_per_block": Tb, "slab_cells": slab_cells, "fields": [f.name for f in fields],
                       "parallelism": "timestep blocks per rank, no collective",
                       "dropouts": ("IDW fields i.i.d. p=0.02; DK block outages quantised to %d masks/year, %d mask(s) in this block"
                                    % (nmask_year, n_dk_masks)) if wl == "c4" else "IDW field i.i.d. p=0.02; gridded fields complete",
                       "l2": "each launch streams %.1f GB of output, far beyond the 126 MB L2 (no flush needed)" % (8e-9 * slab_cells * Tb),
                       "out_dtype": "f64"},
            "per_field_ms": field_ms,
            "per_field_cell_timesteps_per_s": {k: ncell * Tb / (v * 1e-3) for k, v in field_ms.items()},
            "parity_sample": parity,
            "one_off_s": {"inputs": t_inputs, "handle_create": t_create, "dk_weight_build_per_mask_whole_grid": t_dk_build,
                          "grid_local_fit_per_block_host": t_local_fit},
            "dk_weight_build_s": t_dk_build, "dk_masks_in_block": n_dk_masks,
            "dk_stats": {"masks_built": dk_stats[0], "fallback_cells": dk_stats[1], "max_active": dk_stats[2]},
            "value_year_amortised": year,
            "f32_mode": f32, "per_step_call": per_step,
            "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
def roofline(wl, fields, objs, launch_ms, far, far_ms, steps, nslab, slab_cells, Tb, S, dk_stats, hbm_peak, hbm_src, fp64_peak,
             ncell, pts):
    """Per SURVEY.md 8(d): algorithmic bytes per launch / the kernel's measured duration against the measured HBM peak."""
    out_bytes = 8.0 * slab_cells * Tb
    views = {}
    for f in fields:
        if f.method == "idw":
            alg = out_bytes + 8.0 * slab_cells
        elif f.method == "dk":
            alg = out_bytes + slab_cells * (8.0 + dk_stats[2] * 9.0)
        elif f.opts.get("grid_local"):
            alg = out_bytes + slab_cells * (8.0 + 4.0 + 48.0)
        else:
            alg = out_bytes + slab_cells * (8.0 + 4.0 + 24.0)
        ms = launch_ms[f.name]
        views[f.name] = {"method": f.method + ("/grid_local" if f.opts.get("grid_local") else ""), "bound": "hbm",
                         "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak,
                         "algorithmic_bytes_per_launch": alg, "ms_per_launch_all_kernels_of_the_field": ms}
    idw_f = [f for f in fields if f.method == "idw"]
    alg = out_bytes + 8.0 * slab_cells
    if far and far_ms[2] > 0:
        # the dominant kernel, timed alone with CUDA events on the launch stream inside the timed region (three events per
        # launch recorded by the library: smrfb_idw_far_timing); slabs differ in size, so scale to the full-size slab
        n_launch = far_ms[2]
        cells_total = ncell * steps * len(idw_f)
        main_ms = far_ms[1] / n_launch * (slab_cells * n_launch / cells_total)
        node_ms = far_ms[0] / n_launch * (slab_cells * n_launch / cells_total)
        shape = "cells%d_T%d_S%d" % (slab_cells, Tb, S)
        nq = float(np.mean([objs[f.name].kernel_info()["mean_near"] for f in idw_f]))
        # fp64 flop per output: E/O-folded 10-node interpolation of numerator and denominator (40 FMA + 8 add per 4 outputs),
        # 2 FMA per near station, reciprocal (4 FMA) + quotient/retrend (2 FMA), the row contraction of phase A (200 FMA per
        # 128 outputs)
        flops = (22.0 + 4.0 * nq + 12.0 + 3.1) * slab_cells * Tb
        roof = {"kernel": "idw_far_kernel (per IDW field and slab: prep_steps_kernel + idw_far_pack_kernel + idw_far_node_kernel + "
                          "idw_far_kernel; the main kernel is %.0f %% of the field's device time, the node kernel %.0f %%)"
                          % (100 * main_ms / launch_ms[idw_f[0].name], 100 * node_ms / launch_ms[idw_f[0].name]),
                "bound": "hbm", "achieved": alg / (main_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg / (main_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                "traffic": ncu_traffic("idw_far_kernel", shape),
                "traffic_source": "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                  "of this shape); null when this shape was not captured",
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": main_ms, "node_kernel_ms_per_launch": node_ms,
                "launches_timed": int(n_launch),
                "fp64_view": {"achieved_tflops": flops / (main_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                              "frac": flops / (main_ms * 1e-3) / 1e12 / fp64_peak,
                              "flop_per_output": flops / (slab_cells * Tb),
                              "note": "what limits the kernel: ~89 fp64 and ~91 other warp instructions per four outputs at 16 warps per SM and 128 "
                                      "registers; the time follows the fp64 instruction count (DESIGN.md 4.1: tools/dfma_issue_probe.cu, "
                                      "profiles/r02_dfma_issue_probe.txt, profiles/r02_far_variants.txt); the fp64 peak is builder-measured "
                                      "(DFMA issue loop, profiles/r01_peak_probe.json), "
                                      "MEASURED_PEAKS.json holds no fp64 figure"},
                "why": "the reference's dense sum is 4 S = 800 flop per 8-byte output (100 flop/B, far right of the fp64 ridge at 5.2 "
                       "flop/B); the near / far-field split brings it to ~%.0f flop per output so that the kernel comes within reach "
                       "of the HBM write roof" % (flops / (slab_cells * Tb))}
    else:
        f0 = idw_f[0] if idw_f else fields[0]
        roof = dict(views[f0.name])
        roof["kernel"] = "all kernels of field %s (the near / far-field IDW path did not apply)" % f0.name
        roof["traffic"] = None
        roof["peak_source"] = hbm_src
    roof["per_field"] = views
    tot_alg = sum(v["algorithmic_bytes_per_launch"] for v in views.values())
    tot_ms = sum(v["ms_per_launch_all_kernels_of_the_field"] for v in views.values())
    roof["whole_step"] = {"achieved": tot_alg / (tot_ms * 1e-3) / 1e9, "frac": tot_alg / (tot_ms * 1e-3) / 1e9 / hbm_peak, "unit": "GB/s",
                          "note": "all fields of one slab, every kernel included"}
    return roof


def parity_sample(fields, objs, launch, slab, d_out, rows, x, y, dem, sta, pts, nx, Tb, torch, ncells=192, nsteps=6):
    """Re-run every field once on the first slab and compare sampled (cell, step) outputs with the CPU oracle."""
    from oracle import oracle as orc
    import warnings
    rng = np.random.default_rng(99)
    r0, r1 = slab
    ci = rng.integers(r0, r1, ncells)
    cj = rng.integers(0, nx, ncells)
    ts = np.sort(rng.choice(Tb, nsteps, replace=False))
    Xs, Ys, Zs = x[cj][None, :], y[ci][None, :], dem[ci, cj][None, :]
    flat = torch.from_numpy(((ci - r0) * nx + cj).astype(np.int64)).cuda()
    mx, my, mz = sta
    res = {}
    worst_all = 0.0
    for f in fields:
        launch(f, slab)
        torch.cuda.synchronize()
        got = d_out[torch.from_numpy(ts.astype(np.int64)).cuda()][:, flat].cpu().numpy()          # [nsteps, ncells]
        lo, hi = f.vmin, f.vmax
        n = ncells
        if f.method == "idw":
            o = orc.IDW(mx, my, Xs, Ys, mz=mz, GridZ=Zs, power=2.0)
            ref = lambda row: o.detrendedIDW(row.copy(), f.slope) if f.detrend else o.calculateIDW(row.copy())
        elif f.method == "dk":
            n = 24                                              # the CPU kriging solve is ~10 cells/s/core
            impl = "ref" if orc.krige_ref_lib() is not None else "oracle"
            o = orc.DK(mx, my, mz, Xs[:, :n], Ys[:, :n], Zs[:, :n], {"dk_ncores": os.cpu_count() or 1, "detrend_slope": f.slope}, impl=impl)
            ref = lambda row: o.calculate(row.copy())
        else:
            cfg = {"grid_local": bool(f.opts.get("grid_local")), "grid_mask": False, "grid_local_n": f.opts.get("grid_local_n", 25)}
            o = orc.GRID(cfg, pts[0], pts[1], Xs, Ys, mz=pts[2], GridZ=Zs, metadata=f.opts.get("metadata"))
            ref = lambda row: (o.detrendedInterpolation(row.copy(), f.slope, "linear") if f.detrend
                               else o.calculateInterpolation(row.copy(), "linear"))
        worst = 0.0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k, t in enumerate(ts):
                row = rows[f.name][t]
                want = orc.set_min_max(np.asarray(ref(row), dtype=float).reshape(1, -1), lo, hi)[0]
                worst = max(worst, orc.parity_error(got[k, :n], want[:n], float(np.nanmax(np.abs(row)))))
        res[f.name] = {"max_err": worst, "n": int(n * nsteps)}
        worst_all = max(worst_all, worst)
    return {"max_err": worst_all, "n": int(sum(v["n"] for v in res.values())), "tolerance": 1e-9,
            "metric": "|gpu - oracle| / max(|oracle|, max|station data of the step|), sampled (cell, step) outputs of one slab",
            "per_field": res}


def f32_mode(fields, objs, launch, slab, d_out, Tb, nx, L, _lib, torch, stream, reps=3):
    """float32 stores (smrfb_set_output_dtype): device-resident rate of every field on the first slab, 4 bytes per output."""
    out = {}
    cells = (slab[1] - slab[0]) * nx
    for f in fields:
        h = objs[f.name]._h
        _lib.check(L.smrfb_set_output_dtype(h, 1))
        try:
            launch(f, slab)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                launch(f, slab)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[f.name] = {"cell_timesteps_per_s": cells * Tb / (ms * 1e-3), "ms_per_launch": ms,
                           "write_gbs": 4.0 * cells * Tb / (ms * 1e-3) / 1e9}
        finally:
            _lib.check(L.smrfb_set_output_dtype(h, 0))
    tot = sum(v["ms_per_launch"] for v in out.values())
    return {"value": len(out) * cells * Tb / (tot * 1e-3), "unit": "cell-timesteps/s", "out_dtype": "f32", "contract": 1e-5,
            "note": "secondary figure: same fp64 arithmetic, value rounded once at the store (the reference's NetCDF default type)",
            "per_field": out}


def d2h_ceiling(torch, mib=1024):
    """What one rank alone gets out of its PCIe link: device -> pinned host copy of 1 GiB, best of 3."""
    try:
        src = torch.empty(mib << 20, dtype=torch.uint8, device="cuda")
        dst = torch.empty(mib << 20, dtype=torch.uint8, pin_memory=True)
        best = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, (mib << 20) / (time.perf_counter() - t0) / 1e9)
        return best
    except Exception:
        return None


def per_step_call(wl, x, y, dem, sta, pts, series, fields, device, nsteps=3):
    """T = 1 through the reference-shaped image_data._distribute (pandas Series in, [ny, nx] float64 out, clamp fused) on a
    1000 x 1000 window of the basin: what an unmodified smrf.framework run loop pays per variable and timestep."""
    import pandas as pd
    from smrf_b200.distribute.image_data import image_data

    class Topo:
        pass
    n = min(1000, len(y), len(x))
    topo = Topo()
    topo.X, topo.Y = np.meshgrid(x[:n], y[:n])
    topo.dem = np.ascontiguousarray(dem[:n, :n])
    topo.mask = np.ones((n, n))
    f = fields[0]
    if f.method in ("idw", "dk"):
        mx, my, mz = sta
        names = ["s%03d" % i for i in range(len(mx))]
        md = pd.DataFrame({"utm_x": mx, "utm_y": my, "elevation": mz}, index=names)
        cfg = {"distribution": f.method, "detrend": True, "detrend_slope": f.slope, "idw_power": 2.0, "min": f.vmin, "max": f.vmax,
               "stations": None, "dk_ncores": 1}
    else:
        px, py, pz = pts
        names = ["p%04d" % i for i in range(len(px))]
        md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                           "longitude": -116.0 + (px - px.min()) / 82000.0}, index=names)
        cfg = {"distribution": "grid", "detrend": True, "detrend_slope": f.slope, "grid_method": "linear", "grid_local": False,
               "grid_mask": False, "min": f.vmin, "max": f.vmax, "stations": None}
    im = image_data(f.name)
    im.getConfig(cfg)
    im._initialize(topo, md)
    data = pd.DataFrame(series[f.name][:nsteps + 1], columns=names)
    im._distribute(data.iloc[0])
    t0 = time.perf_counter()
    for t in range(1, nsteps + 1):
        im._distribute(data.iloc[t])
    dt = (time.perf_counter() - t0) / nsteps
    return {"field": f.name, "method": f.method, "grid": "%dx%d" % (n, n), "ms_per_call": dt * 1e3,
            "cell_timesteps_per_s": n * n / dt, "api": "smrf_b200.distribute.image_data._distribute(pd.Series) -> [ny, nx] float64, clamp fused"}


if __name__ == "__main__":
    sys.exit(main())
