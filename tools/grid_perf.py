"""Device-resident timing of the GRID-linear kernel on the C5 shape (development aid)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from smrf_b200 import _lib, synthetic as syn
from smrf_b200.spatial import GRID
ny = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
x, y, dem = syn.make_dem(ny, nx, dx=10.0, noise=False)
px, py, pz = syn.make_source_lattice(x, y, dem, spacing=3000.0 * ny / 4000 / 5.0 if ny < 4000 else 600.0, jitter=30.0, seed=5)
X = np.broadcast_to(x[None, :], (ny, nx)); Y = np.broadcast_to(y[:, None], (ny, nx))
t0 = time.perf_counter()
g = GRID({"grid_local": False, "grid_mask": False}, px, py, X, Y, mz=pz, GridZ=dem)
print("points %d, create (Delaunay + find_simplex on host + upload): %.2f s" % (len(px), time.perf_counter() - t0))
f = syn.make_source_fields(px, py, pz, T, seed=9)
L = _lib.lib()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
d_f = torch.from_numpy(f).cuda()
d_out = torch.empty((T, ny * nx), dtype=torch.float64, device="cuda")
for det in (1, 0):
    for _ in range(2):
        _lib.check(L.smrfb_grid_distribute_dev(g._h, d_f.data_ptr(), T, det, -1, 0, -50.0, 50.0, None, None, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        _lib.check(L.smrfb_grid_distribute_dev(g._h, d_f.data_ptr(), T, det, -1, 0, -50.0, 50.0, None, None, d_out.data_ptr(), st))
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    cs = ny * nx * T / (ms * 1e-3)
    print("GRID %dx%d P=%d T=%d detrend=%d: %.2f ms  %.1f G cell-steps/s  write %.0f GB/s" % (ny, nx, len(px), T, det, ms, cs / 1e9, cs * 8 / 1e9))

# grid_local (detrendedInterpolationLocal): three point fields per step, one kernel
import pandas as pd
md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                   "longitude": -116.0 + (px - px.min()) / 82000.0})
t0 = time.perf_counter()
gl = GRID({"grid_local": True, "grid_local_n": 25, "grid_mask": False}, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
print("grid_local create (neighbour table as the reference builds it + Delaunay + upload): %.2f s" % (time.perf_counter() - t0))
t0 = time.perf_counter()
res, slope, icpt = gl.local_fit(f, flag=-1, exact=False)
print("local fits of %d steps x %d points (closed form, host): %.3f s" % (T, len(px), time.perf_counter() - t0))
d_f3 = torch.from_numpy(np.ascontiguousarray(np.stack([res, slope, icpt], axis=1))).cuda()
_lib.check(L.smrfb_grid_set_local_exact(gl._h, int(os.environ.get('GRID_LOCAL_EXACT', '0'))))
for _ in range(2):
    _lib.check(L.smrfb_grid_distribute_local_dev(gl._h, d_f3.data_ptr(), T, -50.0, 50.0, d_out.data_ptr(), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3):
    _lib.check(L.smrfb_grid_distribute_local_dev(gl._h, d_f3.data_ptr(), T, -50.0, 50.0, d_out.data_ptr(), st))
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
cs = ny * nx * T / (ms * 1e-3)
print("GRID local %dx%d P=%d T=%d: %.2f ms  %.1f G cell-steps/s  write %.0f GB/s" % (ny, nx, len(px), T, ms, cs / 1e9, cs * 8 / 1e9))

# grid_method='cubic' (Clough-Tocher): host gradient estimation (scipy) + the nine-weight kernel
for name, obj, nf in (("cubic mask-detrended", g, 1), ("cubic local", gl, 2)):
    t0 = time.perf_counter()
    obj._ensure_cubic()
    t_w = time.perf_counter() - t0
    t0 = time.perf_counter()
    grad = obj._gradients(f if nf == 1 else np.concatenate([res, icpt, slope], axis=0))
    t_g = time.perf_counter() - t0
    fields = np.zeros((T, nf, 3, len(px)))
    if nf == 1:
        fields[:, 0, 0], fields[:, 0, 1], fields[:, 0, 2] = f, grad[:, :, 0], grad[:, :, 1]
        pvs = torch.from_numpy(np.tile(np.array([-0.0065, 20.0]), (T, 1))).cuda()
    else:
        fields[:, 0, 0] = res + icpt
        fields[:, 0, 1], fields[:, 0, 2] = grad[:T, :, 0] + grad[T:2 * T, :, 0], grad[:T, :, 1] + grad[T:2 * T, :, 1]
        fields[:, 1, 0], fields[:, 1, 1], fields[:, 1, 2] = slope, grad[2 * T:, :, 0], grad[2 * T:, :, 1]
        pvs = None
    d_fc = torch.from_numpy(fields).cuda()
    pvp = pvs.data_ptr() if pvs is not None else None
    for _ in range(2):
        _lib.check(L.smrfb_grid_distribute_cubic_dev(obj._h, d_fc.data_ptr(), T, nf, pvp, -50.0, 50.0, d_out.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        _lib.check(L.smrfb_grid_distribute_cubic_dev(obj._h, d_fc.data_ptr(), T, nf, pvp, -50.0, 50.0, d_out.data_ptr(), st))
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    cs = ny * nx * T / (ms * 1e-3)
    print("GRID %s %dx%d P=%d T=%d: %.2f ms  %.1f G cell-steps/s  write %.0f GB/s   (weights once %.2f s, host gradients of %d columns %.2f s)"
          % (name, ny, nx, len(px), T, ms, cs / 1e9, cs * 8 / 1e9, t_w, grad.shape[0], t_g))
