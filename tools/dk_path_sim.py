"""CPU study of the kriging elimination paths (krige.c:134-198) on the C4 geometry: how much of the walk can
neighbouring cells share?  Counts, per tile shape, the distinct path prefixes (tree nodes = rank-1 downdates of the
shared walk in csrc/dk.cu) and the distinct active SETS (what a walk keyed by set instead of by path would do).

    python tools/dk_path_sim.py [row0 col0 [ny nx]]
"""
import sys
import numpy as np

S = 200
dx = 30.0
rng = np.random.default_rng(7)
jj = rng.integers(0, 10000, size=S)
ii = rng.integers(0, 10000, size=S)
x0, y0 = 500000.0, 4800000.0
mx = x0 + dx * jj + 0.37 * dx
my = y0 - dx * ii + 0.37 * dx
mz = (1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0)
      + 120.0 * np.sin(2 * np.pi * mx / 3100.0 + 1.0) * np.sin(2 * np.pi * my / 2300.0)) + rng.normal(0, 5.0, S)

r0 = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
c0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
ny = int(sys.argv[3]) if len(sys.argv) > 3 else 32
nx = int(sys.argv[4]) if len(sys.argv) > 4 else 32

order = np.lexsort((np.arange(S), -mz))          # elevation desc, index asc
sx, sy = mx[order], my[order]
n = S
NN = n + 1
ad = np.sqrt((sx[:, None] - sx[None, :]) ** 2 + (sy[:, None] - sy[None, :]) ** 2)
M = np.zeros((NN, NN))
M[:n, :n] = ad
M[:n, n] = 1.0
M[n, :n] = 1.0
B0 = np.linalg.inv(M)
B0 = 0.5 * (B0 + B0.T)

xs = x0 + dx * (c0 + np.arange(nx))
ys = y0 - dx * (r0 + np.arange(ny))
X, Y = np.meshgrid(xs, ys)
C = ny * nx
rhs = np.ones((C, NN))
rhs[:, :n] = np.sqrt((X.reshape(-1, 1) - sx[None, :]) ** 2 + (Y.reshape(-1, 1) - sy[None, :]) ** 2)

B = np.broadcast_to(B0, (C, NN, NN)).copy()
w = rhs @ B0
act = np.ones((C, n), dtype=bool)
paths = -np.ones((C, n), dtype=np.int32)
alive = np.ones(C, dtype=bool)
for depth in range(n - 1):
    neg = act & (w[:, :n] < 0.0)
    has = neg.any(axis=1) & alive
    if not has.any():
        break
    idx = np.nonzero(has)[0]
    k = neg[idx].argmax(axis=1)
    paths[idx, depth] = k
    Bk = B[idx, :, k]                                   # [m, NN]
    beta = Bk[np.arange(len(idx)), k]
    f = w[idx, k] / beta
    w[idx] -= f[:, None] * Bk
    w[idx, k] = 0.0
    B[idx] -= Bk[:, :, None] * (Bk[:, None, :] / beta[:, None, None])
    act[idx, k] = False
    alive &= has
plen = (paths >= 0).sum(axis=1)
print("region rows %d.. cols %d.. %dx%d: path length mean %.1f, final active mean %.2f max %d"
      % (r0, c0, ny, nx, plen.mean(), act.sum(1).mean(), act.sum(1).max()))
P = paths.reshape(ny, nx, n)


def count(ty, tx):
    tot_tree = tot_set = 0
    ntile = 0
    first_fork = []
    for a in range(0, ny, ty):
        for b in range(0, nx, tx):
            pp = P[a:a + ty, b:b + tx].reshape(-1, n)
            tree = set()
            sets = set()
            for p in pp:
                L = int((p >= 0).sum())
                cur = []
                for d in range(L):
                    cur.append(int(p[d]))
                    tree.add(tuple(cur))
                    sets.add((frozenset(cur), ))
            # depth of the first disagreement
            ff = n
            for d in range(n):
                col = pp[:, d]
                if len(set(col.tolist())) > 1:
                    ff = d
                    break
            first_fork.append(ff)
            tot_tree += len(tree)
            tot_set += len(sets)
            ntile += 1
    cells = ny * nx
    print("tile %3dx%-3d: tree nodes/tile %8.1f (per cell %6.2f)   distinct sets/tile %8.1f (per cell %6.2f)   first fork depth mean %.1f min %d"
          % (ty, tx, tot_tree / ntile, tot_tree / cells, tot_set / ntile, tot_set / cells, np.mean(first_fork), min(first_fork)))


for ty, tx in ((1, 1), (1, 32), (4, 8), (8, 8), (8, 16), (16, 16), (32, 32), (ny, nx)):
    if ty <= ny and tx <= nx:
        count(ty, tx)
