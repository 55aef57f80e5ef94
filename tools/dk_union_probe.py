"""How large is the union of the kriging active sets over 32 consecutive cells of a row (C4 station density)?"""
import sys
import numpy as np
sys.path.insert(0, ".")
from smrf_b200.spatial import DK
ny, nx, S = 128, 2048, 200
rng = np.random.default_rng(7)
dx = 30.0
x = 500000.0 + 141000.0 + dx * np.arange(nx)
y = 4800000.0 - 97000.0 - dx * np.arange(ny)
X, Y = np.meshgrid(x, y)
dem = (1800.0 + 700.0 * np.sin(2 * np.pi * X / 41000.0) * np.cos(2 * np.pi * Y / 57000.0)
       + 120.0 * np.sin(2 * np.pi * X / 3100.0 + 1.0) * np.sin(2 * np.pi * Y / 2300.0)).astype(np.float32).astype(np.float64)
mx = 500000.0 + dx * rng.integers(0, 10000, S) + 0.37 * dx
my = 4800000.0 - dx * rng.integers(0, 10000, S) + 0.37 * dx
mz = 1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0) + rng.normal(0, 5.0, S)
dk = DK(mx, my, mz, X, Y, dem, {"dk_ncores": 1, "detrend_slope": 1})
dk.nan_val = np.zeros(S, dtype=bool)
W = dk.calculateWeights()
W = np.asarray(dk.weights if W is None else W)
nz = W != 0.0
k = nz.sum(axis=2)
print("active per cell: mean %.2f max %d" % (k.mean(), k.max()))
for g in (32, 64):
    u = nz.reshape(ny, nx // g, g, S).any(axis=2).sum(axis=2)
    print("union over %d consecutive cells: mean %.2f  p50 %d p90 %d p99 %d max %d" % (g, u.mean(), np.percentile(u, 50), np.percentile(u, 90), np.percentile(u, 99), u.max()))
u = nz.reshape(ny // 4, 4, nx // 8, 8, S).any(axis=(1, 3)).sum(axis=2)
print("union over 8x4 blocks: mean %.2f p90 %d p99 %d max %d" % (u.mean(), np.percentile(u, 90), np.percentile(u, 99), u.max()))
print(dk.stats())
