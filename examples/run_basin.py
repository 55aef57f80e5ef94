"""A small end-to-end run the way smrf.framework drives this path: the variable classes (air temperature, vapor pressure with
the dew-point consumer, wind with its azimuth, cloud factor) are initialised with a topo and a data object, one thread per
variable distributes hourly station rows (model_framework.py:568-632), the per-step calls are served from fused block launches
(block prefetch) and every grid goes to a float32 NetCDF sink (output_netcdf.py:58,90-93).  Synthetic basin and station series
(smrf_b200.synthetic); needs a CUDA device.

    python examples/run_basin.py [out_dir] [ny nx stations hours block_steps]
"""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smrf_b200 import synthetic as syn  # noqa: E402
from smrf_b200.distribute.air_temp import ta  # noqa: E402
from smrf_b200.distribute.cloud_factor import cf  # noqa: E402
from smrf_b200.distribute.vapor_pressure import vp  # noqa: E402
from smrf_b200.distribute.wind import Wind  # noqa: E402
from smrf_b200.output import BlockSink  # noqa: E402
from smrf_b200.run_loop import run_threaded  # noqa: E402

CFG = {"air_temp": {"distribution": "idw", "detrend": True, "detrend_slope": -1, "idw_power": 2.0, "min": -73.0, "max": 47.0, "stations": None},
       "vapor_pressure": {"distribution": "idw", "detrend": True, "detrend_slope": -1, "idw_power": 2.0, "min": 20.0, "max": 5000.0,
                          "stations": None, "dew_point_tolerance": 0.01, "dew_point_nthreads": 2},
       "wind": {"distribution": "idw", "detrend": False, "detrend_slope": 0, "idw_power": 2.0, "min": 0.447, "max": 35.0, "stations": None},
       "cloud_factor": {"distribution": "idw", "detrend": False, "detrend_slope": 0, "idw_power": 2.0, "min": 0.0, "max": 1.0, "stations": None}}


class Topo:
    pass


class Data:
    pass


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "run_basin_out"
    ny, nx, S, T, nb = (int(a) for a in sys.argv[2:7]) if len(sys.argv) > 6 else (400, 600, 24, 48, 24)
    x, y, dem = syn.make_dem(ny, nx, dx=30.0, seed=3)
    topo = Topo()
    topo.X, topo.Y = np.meshgrid(x, y)
    topo.dem = dem
    topo.mask = np.ones_like(dem)
    mx, my, mz = syn.make_stations(S, x, y, dem, seed=4)
    names = ["ST%02d" % i for i in range(S)]
    data = Data()
    data.metadata = pd.DataFrame({"utm_x": mx, "utm_y": my, "elevation": mz}, index=names)
    times = pd.date_range("1998-01-14 15:00", periods=T, freq="h")
    rng = np.random.default_rng(3)

    def frame(kind, seed):
        return pd.DataFrame(syn.dropouts_iid(syn.make_series(kind, mz, T, seed=seed), p=0.05, seed=seed + 50), index=times, columns=names)
    frames = {"air_temp": frame("air_temp", 3), "vapor_pressure": frame("vapor_pressure", 4),
              "wind_speed": pd.DataFrame(np.abs(rng.normal(4.0, 3.0, (T, S))), index=times, columns=names),
              "wind_direction": pd.DataFrame(rng.uniform(0.0, 360.0, (T, S)), index=times, columns=names),
              "cloud_factor": pd.DataFrame(rng.uniform(-0.1, 1.1, (T, S)), index=times, columns=names)}
    variables = {"air_temp": ta(dict(CFG["air_temp"])), "vapor_pressure": vp(dict(CFG["vapor_pressure"])),
                 "wind": Wind(dict(CFG["wind"])), "cloud_factor": cf(dict(CFG["cloud_factor"]))}
    for v in variables.values():
        v.initialize(topo, data, times)
    outputs = ["air_temp", "vapor_pressure", "dew_point", "wind_speed", "wind_direction", "cloud_factor"]
    sink = BlockSink(out_dir, outputs, x, y, hours=np.arange(T), start_date=str(times[0]), max_pending=8)
    index = {t: i for i, t in enumerate(times)}
    t0 = time.perf_counter()
    run_threaded(variables, frames, times, outputs=outputs, block_steps=nb,
                 sink=lambda n, t, grid: sink(n, (index[t], index[t] + 1), (0, grid.shape[0]), grid[None].copy()))
    sink.close()
    dt = time.perf_counter() - t0
    print("%d x %d cells, %d stations, %d hours, %d fields -> %s: %.2f s (%.1f M cell-timesteps/s incl. the NetCDF files)"
          % (ny, nx, S, T, len(outputs), out_dir, dt, ny * nx * T * len(outputs) / dt / 1e6))


if __name__ == "__main__":
    main()
