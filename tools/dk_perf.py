"""Timing of the DK weight build + steady state on a sub-grid of the C4 basin (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
os.environ["SMRFB_DK_VERBOSE"] = "1"
from smrf_b200 import synthetic as syn
from smrf_b200.spatial import DK
ny = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
x, y, dem = syn.make_dem(10000, 10000, dx=30.0, noise=False)
mx, my, mz = syn.make_stations(S, x, y, dem, seed=7)
r0 = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
c0 = int(sys.argv[5]) if len(sys.argv) > 5 else 4000
xs, ys, ds = x[c0:c0 + nx], y[r0:r0 + ny], np.ascontiguousarray(dem[r0:r0 + ny, c0:c0 + nx])
X = np.broadcast_to(xs[None, :], (ny, nx)); Y = np.broadcast_to(ys[:, None], (ny, nx))
dk = DK(mx, my, mz, X, Y, ds, {"detrend_slope": 1})
data = np.abs(syn.make_series("air_temp", mz, 8, seed=1))
t0 = time.perf_counter(); dk.distribute_block(data[:1]); t1 = time.perf_counter()
print("build+1 step: %.2f s  (%.0f cells/s)" % (t1 - t0, ny * nx / (t1 - t0)), dk.stats())
t0 = time.perf_counter(); dk.distribute_block(data); t1 = time.perf_counter()
print("8 steps host API: %.3f s" % (t1 - t0))
for rep in range(int(os.environ.get('DK_PERF_REBUILDS', '2'))):          # rebuild with other masks: kernel-time repeatability
    d2 = data[:1].copy(); d2[0, rep] = np.nan
    t0 = time.perf_counter(); dk.distribute_block(d2); t1 = time.perf_counter()
    print("rebuild %d: %.2f s" % (rep, t1 - t0))
