#!/usr/bin/env python
"""bench.py -- distributed cell-timesteps/sec of the SMRF spatial-distribution hot path on B200.

Workload (BASELINE.json configs[3], "C4"): synthetic 10 000 x 10 000 DEM (1e8 cells, 30 m), 200
stations, hourly series of one year; one *step* = one block of T_b = 720 hourly timesteps (30 days)
distributed for three fields over the whole grid, walked slab by slab (1000 DEM rows = 1e7 cells per
slab) so that the device-resident output block [T_b][slab] stays at 58 GB:
    air_temp        IDW, detrended (slope -1), clamp [-73, 47], i.i.d. 2 % station dropouts
    vapor_pressure  IDW, detrended (slope -1), clamp [20, 5000], i.i.d. 2 % station dropouts
    precip          DK,  detrended (slope +1), clamp [0, inf), block outages (<= 64 masks / year)
value = cells * T_b * 3 fields * steps / time, inputs resident in HBM, output into a
device-resident block (CUDA events on the launch stream, max over ranks).  DK kriging weights are
a per-NaN-mask one-off: they are built during warm-up and reported separately (dk_weight_build_s).
e2e   = the same fields through the public Python classes (smrf_b200.spatial.IDW / DK
.distribute_block) with HOST buffers: station rows H2D and the distributed grids D2H into pinned
memory inside the timed region (a shorter block, T_e2e steps, because PCIe bounds it).
Multi-GPU (torchrun, one rank per GPU): timestep blocks shard across ranks (rank r distributes
hours [r*T_b, (r+1)*T_b)), no collective on the data path; weak scaling.

--impl reference times the reference's CPU path (NumPy IDW as idw.py does it + the reference's
krige.c compiled in place when oracle/_ref is present, else the C restatement) on a bounded tile.
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

FIELDS = (
    # name, method, slope flag, vmin, vmax
    ("air_temp", "idw", -1, -73.0, 47.0),
    ("vapor_pressure", "idw", -1, 20.0, 5000.0),
    ("precip", "dk", 1, 0.0, float("inf")),
)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ny", type=int, default=10000)
    ap.add_argument("--nx", type=int, default=10000)
    ap.add_argument("--stations", type=int, default=200)
    ap.add_argument("--tb", type=int, default=720, help="timesteps per block (30 days of hourly steps)")
    ap.add_argument("--slab-rows", type=int, default=1000, help="DEM rows per cell slab; the device-resident output block is [tb][slab]")
    ap.add_argument("--te2e", type=int, default=0,
                    help="timesteps per block on the host-buffer (e2e) leg; 0 = max(8, 32 // n_gpus): 25.6 GB of pinned "
                         "host memory at 1 GPU, 6.4 GB per rank at 4-8 GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic basin
# ------------------------------------------------------------------------------------------------
def make_inputs(ny, nx, S, T_total):
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(ny, nx, dx=30.0, noise=(ny * nx <= 25_000_000))
    if ny * nx > 25_000_000:      # cheap deterministic micro-relief instead of 1e8 normal draws
        j = np.arange(nx, dtype=np.float64)[None, :]
        i = np.arange(ny, dtype=np.float64)[:, None]
        dem = (dem + 8.0 * np.sin(0.7 * j + 1.3 * i) * np.cos(0.31 * i - 0.2 * j)).astype(np.float32).astype(np.float64)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
    series = {}
    series["air_temp"] = syn.dropouts_iid(syn.make_series("air_temp", mz, T_total, seed=3), p=0.02, seed=11)
    series["vapor_pressure"] = syn.dropouts_iid(syn.make_series("vapor_pressure", mz, T_total, seed=4), p=0.02, seed=12)
    pr = syn.make_series("precip", mz, T_total, seed=5)
    pr, nmask = syn.dropouts_block(pr, max_masks=12, seed=13)   # <= 64 allowed; 12 => one mask per 730 h
    series["precip"] = pr
    return x, y, dem, mx, my, mz, series, nmask


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on a bounded tile
# ------------------------------------------------------------------------------------------------
def cpu_arm(S, steps, warmup, idw_tile=(400, 400), dk_tile=(24, 24), t_per_step=2):
    from oracle import oracle as orc
    from oracle import build as obuild
    from smrf_b200 import synthetic as syn
    obuild.build_oracle()
    kind = "reference" if orc.krige_ref_lib() is not None else "port"
    impl = "ref" if kind == "reference" else "oracle"
    ncores = os.cpu_count() or 1
    x, y, dem = syn.make_dem(2000, 2000, dx=30.0, noise=False)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
    T = (steps + warmup) * t_per_step
    ta = syn.dropouts_iid(syn.make_series("air_temp", mz, T, seed=3), p=0.02, seed=11)
    vp = syn.dropouts_iid(syn.make_series("vapor_pressure", mz, T, seed=4), p=0.02, seed=12)
    pr = syn.make_series("precip", mz, 400, seed=5)
    pr = pr[pr.sum(axis=1) > 0][:T]
    while len(pr) < T:
        pr = np.concatenate([pr, pr])[:T]
    iy, ix = idw_tile
    Xi, Yi = np.meshgrid(x[800:800 + ix], y[700:700 + iy])
    Zi = dem[700:700 + iy, 800:800 + ix]
    ky, kx = dk_tile
    Xk, Yk = np.meshgrid(x[800:800 + kx], y[700:700 + ky])
    Zk = dem[700:700 + ky, 800:800 + kx]
    t0 = time.perf_counter()
    idw = orc.IDW(mx, my, Xi, Yi, mz=mz, GridZ=Zi, power=2.0)          # the [ny,nx,S] cubes, as idw.py:53-77
    t_idw_setup = time.perf_counter() - t0
    dk = orc.DK(mx, my, mz, Xk, Yk, Zk, {"dk_ncores": ncores, "detrend_slope": 1}, impl=impl)
    t0 = time.perf_counter()
    dk.calculate(pr[0].copy())                                          # builds the kriging weights (one mask)
    t_dk_build = time.perf_counter() - t0

    def one_step(k):
        # one Python thread per variable, as smrf/framework/model_framework.py:615-632
        def f_idw(rows, lo, hi):
            for r in rows:
                orc.set_min_max(idw.detrendedIDW(r.copy(), -1), lo, hi)

        def f_dk(rows):
            for r in rows:
                orc.set_min_max(dk.calculate(r.copy()), 0.0, None)
        sl = slice(k * t_per_step, (k + 1) * t_per_step)
        th = [threading.Thread(target=f_idw, args=(ta[sl], -73.0, 47.0)),
              threading.Thread(target=f_idw, args=(vp[sl], 20.0, 5000.0)),
              threading.Thread(target=f_dk, args=(pr[sl],))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for k in range(warmup):
        one_step(k)
    times = [one_step(warmup + k) for k in range(steps)]
    tot = sum(times)
    # per step: 2 IDW fields on the IDW tile + 1 DK field on the DK tile, run concurrently
    cell_steps = steps * t_per_step * (2 * iy * ix + ky * kx)
    value = cell_steps / tot
    sample = ("%d steps x %d hours: IDW air_temp+vapor_pressure on a %dx%d tile + DK precip on a %dx%d tile, %d stations, "
              "3 Python threads (one per variable) + %d OpenMP threads for krige; NumPy cubes as idw.py builds them; "
              "DK weight build (one mask, %d cells) took %.1f s = %.1f cells/s and is excluded like on the GPU; IDW cube setup %.1f s"
              % (steps, t_per_step, iy, ix, ky, kx, S, ncores, ky * kx, t_dk_build, ky * kx / t_dk_build, t_idw_setup))
    return {"value": value, "unit": "cell-timesteps/s", "cores": ncores, "kind": kind, "sample": sample,
            "dk_weight_build_cells_per_s": ky * kx / t_dk_build}, tot / steps * 1e3


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


I8_TRAFFIC = 57.64e9  # dram bytes of one idw_i8_kernel launch at the C4 slab shape (ncu --set full, profiles/r01_i8_ncu.md)


def load_peaks():
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm = float(json.load(open(p))["hbm_gbs"])
            src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    fp64, fsrc = 37.18, "measured (profiles/r01_peak_probe.json, DMMA m8n8k4 issue loop)"
    q = os.path.join(ROOT, "profiles", "r01_peak_probe.json")
    if os.path.exists(q):
        try:
            fp64 = float(json.load(open(q))["fp64_dmma_m8n8k4_tflops"])
        except Exception:
            pass
    return hbm, src, fp64, fsrc


def int8_peak_tops(sm_mhz):
    """Dense int8 tensor peak of this GPU in TOP/s: tcgen05.mma kind::i8 issue loop measured by tools/umma_i8_probe.cu
    (profiles/r01_umma_i8_probe.json: MAC per clock and SM) x 2 x SMs x the SM clock."""
    mac, sms = 8187.0, 148
    q = os.path.join(ROOT, "profiles", "r01_umma_i8_probe.json")
    if os.path.exists(q):
        try:
            d = json.load(open(q))
            mac, sms = float(d["int8_mac_per_clk_per_sm"]), int(d["sms"])
        except Exception:
            pass
    return 2.0 * mac * sms * sm_mhz * 1e6 / 1e12


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, ms = cpu_arm(args.stations, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "distributed cell-timesteps/sec (IDW+DK)", "value": cb["value"],
                "unit": "cell-timesteps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "C4 synthetic basin (BASELINE.json configs[3]): IDW air_temp + vapor_pressure, DK precip, "
                                       "%d stations; CPU tile sample of it (the reference's [ny,nx,S] cubes need 320 GB at 1e8 cells)"
                                       % args.stations},
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
    from smrf_b200.spatial import DK, IDW
    L = _lib.lib()

    ny, nx, S, Tb = args.ny, args.nx, args.stations, args.tb
    ncell = ny * nx
    T_total = 8760
    x, y, dem, mx, my, mz, series, nmask_year = make_inputs(ny, nx, S, T_total)
    # GridX / GridY as np.meshgrid(x, y) would give them, but as zero-copy broadcast views (1.6 GB saved at 1e8 cells)
    X = np.broadcast_to(x[None, :], (ny, nx))
    Y = np.broadcast_to(y[:, None], (ny, nx))
    seg = T_total // 12                  # DK outage masks change every 730 h: block r sits inside segment r
    t_lo = (rank * seg) % (T_total - Tb)
    rows = {k: np.ascontiguousarray(v[t_lo:t_lo + Tb]) for k, v in series.items()}
    # precip rows that are entirely zero are skipped by the reference (precipitation.py:335); keep them: the
    # distribution cost does not depend on the values
    objs = {}
    t0 = time.perf_counter()
    for name, method, flag, lo, hi in FIELDS:
        if method == "idw":
            objs[name] = IDW(mx, my, X, Y, mz=mz, GridZ=dem, power=2.0, device=local_rank)
        else:
            objs[name] = DK(mx, my, mz, X, Y, dem, {"dk_ncores": 1, "detrend_slope": flag}, device=local_rank)
    t_create = time.perf_counter() - t0
    del X, Y

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    d_rows = {k: torch.from_numpy(v).cuda() for k, v in rows.items()}
    slab_rows = min(args.slab_rows, ny)
    slabs = [(r0 * nx, (min(r0 + slab_rows, ny) - r0) * nx) for r0 in range(0, ny, slab_rows)]
    slab_cells = max(c for _, c in slabs)
    d_out = torch.empty((Tb, slab_cells), dtype=torch.float64, device="cuda")
    n_dk_masks = len({np.isnan(r).tobytes() for r in rows["precip"]})

    ev_field = {}

    def device_step(record=None):
        for cbeg, ccnt in slabs:
            for name, method, flag, lo, hi in FIELDS:
                if record is not None:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                if method == "idw":
                    _lib.check(L.smrfb_idw_distribute_window_dev(objs[name]._h, d_rows[name].data_ptr(), Tb, 1, flag, lo, hi,
                                                                 None, None, cbeg, ccnt, d_out.data_ptr(), st))
                else:
                    _lib.check(L.smrfb_dk_distribute_window_dev(objs[name]._h, d_rows[name].data_ptr(), _lib.dptr(rows[name]),
                                                                Tb, lo, hi, None, None, cbeg, ccnt, d_out.data_ptr(), st))
                if record is not None:
                    e1.record(stream)
                    record.setdefault(name, []).append((e0, e1, ccnt))

    # DK weights: one-off per NaN mask, built on the first call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.smrfb_dk_distribute_window_dev(objs["precip"]._h, d_rows["precip"].data_ptr(), _lib.dptr(rows["precip"]), Tb,
                                                0.0, float("inf"), None, None, slabs[0][0], slabs[0][1], d_out.data_ptr(), st))
    torch.cuda.synchronize()
    t_dk_build = time.perf_counter() - t0
    dk_stats = objs["precip"].stats()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
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
    tms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    ms_step = ms_total / args.steps
    value = world * ncell * Tb * len(FIELDS) * args.steps / (ms_total * 1e-3)
    # per field: ms for one pass over the whole grid (all slabs), and per launch (one slab)
    field_ms = {k: float(sum(a.elapsed_time(b) for a, b, _ in v)) / args.steps for k, v in ev_field.items()}
    launch_ms = {k: float(np.mean([a.elapsed_time(b) for a, b, c in v if c == slab_cells])) for k, v in ev_field.items()}

    # ---- e2e: public Python API, host buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        Te = min(args.te2e if args.te2e > 0 else max(8, 32 // world), Tb)
        del d_out
        torch.cuda.empty_cache()          # the host-buffer API keeps its own device ring per handle
        host_out = torch.empty((Te, ny, nx), dtype=torch.float64, pin_memory=True).numpy()
        host_rows = {k: torch.from_numpy(v[:Te].copy()).pin_memory().numpy() for k, v in rows.items()}

        def e2e_step():
            for name, method, flag, lo, hi in FIELDS:
                if method == "idw":
                    objs[name].distribute_block(host_rows[name], detrend=True, flag=flag, vmin=lo, vmax=hi, out=host_out)
                else:
                    objs[name].distribute_block(host_rows[name], vmin=lo, vmax=None if hi == float("inf") else hi, out=host_out)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * ncell * Te * len(FIELDS) * n_e2e / dt, "unit": "cell-timesteps/s",
               "h2d_bytes_per_step": int(len(FIELDS) * Te * S * 8), "d2h_bytes_per_step": int(len(FIELDS) * Te * ncell * 8),
               "steps_per_block": Te, "blocks_timed": n_e2e,
               "api": "smrf_b200.spatial.IDW/DK.distribute_block(host rows) -> pinned float64 [T, ny, nx]"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, hbm_src, fp64_peak, fp64_src = load_peaks()
    idw_ms = float(np.mean([launch_ms["air_temp"], launch_ms["vapor_pressure"]]))
    idw_flops = 2.0 * S * slab_cells * Tb     # numerator contraction only (the masked denominator is a sparse correction)
    idw_bytes = 8.0 * slab_cells * Tb + 8.0 * slab_cells
    dk_bytes = 8.0 * slab_cells * Tb + slab_cells * (8.0 + dk_stats[2] * 9.0)
    # which IDW kernel ran: batches of >= 32 steps and <= 224 stations take the int8-sliced tcgen05 kernel (idw_i8.cu)
    i8 = (Tb >= 32 and 16 < S <= 224 and os.environ.get("SMRFB_IDW_KERNEL", "") != "dmma")
    common = {"traffic_unit": "bytes per launch (algorithmic: %.3e)" % idw_bytes, "ms_per_launch": idw_ms,
              "hbm_view": {"bound": "hbm", "achieved": idw_bytes / (idw_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": idw_bytes / (idw_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                           "algorithmic_bytes_per_launch": idw_bytes},
              "dk_distribute_kernel": {"bound": "hbm", "achieved": dk_bytes / (launch_ms["precip"] * 1e-3) / 1e9, "peak": hbm_peak,
                                       "unit": "GB/s", "frac": dk_bytes / (launch_ms["precip"] * 1e-3) / 1e9 / hbm_peak,
                                       "algorithmic_bytes_per_launch": dk_bytes, "ms_per_launch": launch_ms["precip"]}}
    c4_shape = (slab_cells == 10_000_000 and Tb == 720 and S == 200)
    if i8:
        nm = min(len(np.unique(np.isnan(rows[f]), axis=0)) for f in ("air_temp", "vapor_pressure"))
        products = 27 + (6 if nm > 32 else 0)        # digit products per output: numerator 27, per-step denominator 6
        KB = (S + 31) // 32
        i8_ops = 2.0 * products * 32 * KB * slab_cells * (-(-Tb // 48) * 48)     # executed int8 ops per launch
        sm_mhz = float((clocks or {}).get("sm_max_mhz") or 1965.0)
        peak = int8_peak_tops(sm_mhz)
        roof = {"kernel": "idw_i8_kernel (prep_steps_kernel + idw_i8_mask_kernel + idw_i8_slice_kernel + idw_i8_kernel per field; "
                          "the three preparation launches are < 0.1 % of it)",
                "bound": "tensor", "achieved": i8_ops / (idw_ms * 1e-3) / 1e12, "peak": peak, "unit": "TOP/s (int8)",
                "frac": i8_ops / (idw_ms * 1e-3) / 1e12 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this exact shape (ncu --set full,
                # profiles/r01_i8_ncu.md); null for other shapes
                "traffic": I8_TRAFFIC if c4_shape else None,
                "peak_source": "int8 tensor peak measured with a tcgen05.mma kind::i8 issue loop (tools/umma_i8_probe.cu, "
                               "profiles/r01_umma_i8_probe.json: 8187 MAC/clk/SM) at the SM clock; MEASURED_PEAKS.json holds bf16 only "
                               "(its dense bf16 figure x 2 = %.0f TOP/s)" % (2 * 1619.0),
                "algorithmic_int8_ops_per_launch": i8_ops, "digit_products_per_output": products, "distinct_nan_masks": int(nm),
                "fp64_equivalent": {"achieved": idw_flops / (idw_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                    "vs_fp64_tensor_peak": idw_flops / (idw_ms * 1e-3) / 1e12 / fp64_peak,
                                    "note": "2*S flop per output as if evaluated in fp64; the DMMA kernel this one replaces reached "
                                            "0.66 of the %.1f TFLOP/s fp64 tensor peak" % fp64_peak},
                "why_tensor": "2*S = 400 fp64-equivalent flop per 8-byte output; evaluated as 27-33 int8 digit-product GEMMs "
                              "(%d int8 op per output) on the 5th-generation tensor cores, s32 accumulators in tensor memory"
                              % int(2 * products * 32 * KB)}
    else:            # fp64 tensor-core (DMMA) kernel: short batches, > 224 stations, or SMRFB_IDW_KERNEL=dmma
        roof = {"kernel": "idw_distribute_kernel (prep_steps_kernel + idw_distribute_kernel per field; the prep launch is < 0.1 % of it)",
                "bound": "tensor", "achieved": idw_flops / (idw_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": idw_flops / (idw_ms * 1e-3) / 1e12 / fp64_peak,
                "traffic": 57.68e9 if c4_shape else None,
                "peak_source": "fp64 tensor (DMMA) peak " + fp64_src + "; MEASURED_PEAKS.json holds no fp64 figure",
                "algorithmic_flops_per_launch": idw_flops,
                "why_tensor": "dense fp64 IDW at 200 stations is 2*S = 400 flop per 8-byte output = 50 flop/B, far right of the fp64 ridge "
                              "(37 TFLOP/s / 6.5 TB/s = 5.7 flop/B): the fp64 tensor pipe, not HBM, bounds it"}
    roof.update(common)
    cb = None
    if not args.no_cpu_baseline and world == 1:      # reported baseline, rank 0 at N = 1 only
        cb, _ = cpu_arm(S, 2, 1)
    line = {"metric": "distributed cell-timesteps/sec (IDW+DK)", "value": value, "unit": "cell-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "arithmetic": ("IDW: fp64 contraction evaluated in integers -- 8-bit digit slices (6 per weight, 7 per residual) on tcgen05.mma kind::i8, "
                           "s32 accumulate, fp64 epilogue; within 1e-11 of the fp64 tensor-core kernel, error bound 2.2e-10 (DESIGN.md 4.1). "
                           "DK, epilogues: fp64"
                           if i8 else "fp64 throughout (DMMA m8n8k4 for the IDW contraction)"),
            "config": {"workload": "C4 (BASELINE.json configs[3]): %dx%d DEM = %d cells, %d stations, hourly; step = %d-hour block x 3 fields "
                                   "(air_temp IDW, vapor_pressure IDW, precip DK) over the whole grid, walked in %d slabs of %d cells"
                                   % (ny, nx, ncell, S, Tb, len(slabs), slab_cells),
                       "timesteps_per_block": Tb, "slab_cells": slab_cells, "fields": [f[0] for f in FIELDS], "parallelism": "timestep blocks per rank, no collective",
                       "dropouts": "IDW fields i.i.d. p=0.02; DK block outages quantised to %d masks/year, %d mask(s) in this block" % (nmask_year, n_dk_masks),
                       "l2": "each launch streams %.1f GB of output, far beyond the 126 MB L2 (no flush needed)" % (8e-9 * slab_cells * Tb),
                       "out_dtype": "f64"},
            "per_field_ms": field_ms,
            "per_field_cell_timesteps_per_s": {k: ncell * Tb / (v * 1e-3) for k, v in field_ms.items()},
            "dk_weight_build_s": t_dk_build, "dk_masks_in_block": n_dk_masks,
            "dk_stats": {"masks_built": dk_stats[0], "fallback_cells": dk_stats[1], "max_active": dk_stats[2]},
            "handle_create_s": t_create,
            "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
