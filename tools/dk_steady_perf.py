"""Device-resident timing of the DK steady-state kernel on a C4 slab shape (development aid)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from smrf_b200 import _lib, synthetic as syn
from smrf_b200.spatial import DK
ny, nx, T = 300, 10000, 720
x, y, dem = syn.make_dem(10000, 10000, dx=30.0, noise=False)
mx, my, mz = syn.make_stations(200, x, y, dem, seed=7)
ys, ds = y[4000:4000 + ny], np.ascontiguousarray(dem[4000:4000 + ny])
X = np.broadcast_to(x[None, :], (ny, nx)); Y = np.broadcast_to(ys[:, None], (ny, nx))
dk = DK(mx, my, mz, X, Y, ds, {"detrend_slope": 1})
data = np.abs(syn.make_series("air_temp", mz, T, seed=1))
L = _lib.lib()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
d_data = torch.from_numpy(data).cuda()
d_out = torch.empty((T, ny * nx), dtype=torch.float64, device="cuda")
def run():
    _lib.check(L.smrfb_dk_distribute_dev(dk._h, d_data.data_ptr(), _lib.dptr(data), T, 0.0, float("inf"), None, None, d_out.data_ptr(), st))
run(); run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): run()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
cs = ny * nx * T / (ms * 1e-3)
print("DK steady %dx%d T=%d K=%d: %.2f ms  %.1f G cell-steps/s  write %.0f GB/s" % (ny, nx, T, dk.stats()[2], ms, cs / 1e9, cs * 8 / 1e9))
