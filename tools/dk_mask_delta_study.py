"""CPU study for the next step of the DK weight build: when ONE station drops out of the mask, for which cells does the
reference's elimination walk (krige.c:134-198) end in a different active set?  If only the cells that USED the station change,
a new mask costs a rebuild of ~K/n of the grid instead of all of it (the weights depend on the final set only).
Oracle: oracle/krige_oracle.c (bit-identical to the reference C).  C4 geometry: 200 stations over 300 km, scattered cells.

    python tools/dk_mask_delta_study.py [ncells nremoved threads]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as orc  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
nrem = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nthreads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
S, dx = 200, 30.0
rng = np.random.default_rng(7)
jj, ii = rng.integers(0, 10000, size=S), rng.integers(0, 10000, size=S)
x0, y0 = 500000.0, 4800000.0
mx, my = x0 + dx * jj + 0.37 * dx, y0 - dx * ii + 0.37 * dx
mz = (1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0)
      + 120.0 * np.sin(2 * np.pi * mx / 3100.0 + 1.0) * np.sin(2 * np.pi * my / 2300.0)) + rng.normal(0, 5.0, S)
crng = np.random.default_rng(99)
cx = x0 + dx * crng.integers(0, 10000, ncell)
cy = y0 - dx * crng.integers(0, 10000, ncell)


def solve(keep):
    """final active flags [ncell, S] (0 for stations outside `keep`) and weights under the mask `keep`"""
    k = np.flatnonzero(keep)
    ad = orc.station_distances(mx[k], my[k])
    dg = np.sqrt((cx[:, None] - mx[k][None, :]) ** 2 + (cy[:, None] - my[k][None, :]) ** 2)
    w = np.zeros((ncell, len(k)))
    act = np.zeros((ncell, len(k)), dtype=np.int32)
    orc.call_grid(ad, dg, mz[k], w, nthreads=nthreads, active=act)
    A = np.zeros((ncell, S), dtype=bool)
    A[:, k] = act != 0
    W = np.zeros((ncell, S))
    W[:, k] = w
    return A, W


t0 = time.perf_counter()
full = np.ones(S, dtype=bool)
A0, W0 = solve(full)
print("all %d stations: %d cells in %.1f s; final sets of %.1f stations on average (max %d)"
      % (S, ncell, time.perf_counter() - t0, A0.sum(1).mean(), A0.sum(1).max()))
use_count = A0.sum(0)
order = np.argsort(-use_count)
picks = list(order[:nrem // 2]) + list(crng.choice(S, nrem - nrem // 2, replace=False))      # heavily used ones and random ones
for s in picks:
    keep = full.copy()
    keep[s] = False
    A1, W1 = solve(keep)
    used = A0[:, s]
    same_set = (A1 == (A0 & keep[None, :])).all(1)
    unused_changed = int((~used & ~same_set).sum())
    unused_w_equal = bool(np.array_equal(W1[~used & same_set], W0[~used & same_set]))
    print("station %3d out: used by %4d of %d cells; of the %4d cells that did not use it %d end in a different set; "
          "weights of the unchanged sets bit-identical: %s" % (s, int(used.sum()), ncell, int((~used).sum()), unused_changed, unused_w_equal))
print("total %.1f s" % (time.perf_counter() - t0))
