"""CPU model of the 'main chain + per-cell extra set' kriging walk (design study for csrc/dk.cu).

A tile keeps ONE inverse S for the set A_main = all stations minus those that EVERY cell of the tile has dropped;
a cell that has also dropped the stations Q (not yet common) gets its weights from the Schur complement
    w = w_main - S[:, Q] (S[Q, Q])^-1 w_main[Q]
so cells step independently (no fork / snapshot) and S is downdated once per station and tile.
Prints the distribution of |Q|, the number of rounds, and compares the final sets with the plain per-cell walk.

    python tools/dk_main_chain_sim.py row0 col0 [ty tx [ntiles_y ntiles_x]]
"""
import sys
import numpy as np

S_ = 200
dx = 30.0
rng = np.random.default_rng(7)
jj = rng.integers(0, 10000, size=S_)
ii = rng.integers(0, 10000, size=S_)
x0, y0 = 500000.0, 4800000.0
mx = x0 + dx * jj + 0.37 * dx
my = y0 - dx * ii + 0.37 * dx
mz = (1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0)
      + 120.0 * np.sin(2 * np.pi * mx / 3100.0 + 1.0) * np.sin(2 * np.pi * my / 2300.0)) + rng.normal(0, 5.0, S_)

r0 = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
c0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
TY = int(sys.argv[3]) if len(sys.argv) > 3 else 4
TX = int(sys.argv[4]) if len(sys.argv) > 4 else 8
NTY = int(sys.argv[5]) if len(sys.argv) > 5 else 2
NTX = int(sys.argv[6]) if len(sys.argv) > 6 else 2
RA = int(sys.argv[7]) if len(sys.argv) > 7 else 4
QMAX = int(sys.argv[8]) if len(sys.argv) > 8 else 12
HANDOFF = 10

order = np.lexsort((np.arange(S_), -mz))
sx, sy = mx[order], my[order]
n = S_
NN = n + 1
ad = np.sqrt((sx[:, None] - sx[None, :]) ** 2 + (sy[:, None] - sy[None, :]) ** 2)
M = np.zeros((NN, NN))
M[:n, :n] = ad
M[:n, n] = 1.0
M[n, :n] = 1.0
B0 = np.linalg.inv(M)
B0 = 0.5 * (B0 + B0.T)


def plain_walk(rhs):
    """reference-like sequential walk of one cell with its own inverse; returns the set at hand-off / finish"""
    B = B0.copy()
    w = B @ rhs
    act = np.ones(n, dtype=bool)
    nd = 0
    while True:
        neg = act & (w[:n] < 0.0)
        if not neg.any():
            return frozenset(np.nonzero(act)[0].tolist()), False
        if n - nd <= HANDOFF:
            return frozenset(np.nonzero(act)[0].tolist()), True
        k = int(neg.argmax())
        b = B[:, k].copy()
        beta = b[k]
        w -= (w[k] / beta) * b
        w[k] = 0.0
        B -= np.outer(b, b / beta)
        B[k, :] = 0.0
        B[:, k] = 0.0
        act[k] = False
        nd += 1


def tile_walk(rhs_list, stats):
    """jobs: (cells, S, mact, nmain); a stuck job (nobody can step, nothing common) is split by the station most
    cells have dropped: those cells go on (progress guaranteed), the others restart later from a snapshot of S"""
    C = len(rhs_list)
    wm = [B0 @ r for r in rhs_list]
    Q = [[] for _ in range(C)]
    done = [False] * C
    hand = [False] * C
    final = [None] * C
    rounds = 0
    jobs = [(list(range(C)), B0.copy(), np.ones(n, dtype=bool), 0)]
    while jobs:
        cells, S, mact, nmain = jobs.pop()
        while True:
            cells = [c for c in cells if not done[c]]
            if not cells:
                break
            rounds += 1
            stepped = 0
            for c in cells:
                for _ in range(RA):
                    q = Q[c]
                    if q:
                        G = S[np.ix_(q, q)]
                        z = np.linalg.solve(G, wm[c][q])
                        w = wm[c] - S[:, q] @ z
                        w[q] = 0.0
                    else:
                        w = wm[c]
                    act = mact.copy()
                    act[q] = False
                    neg = act & (w[:n] < 0.0)
                    if not neg.any() or n - (nmain + len(q)) <= HANDOFF:
                        done[c] = True
                        hand[c] = bool(neg.any())
                        final[c] = frozenset(np.nonzero(act)[0].tolist())
                        break
                    if len(q) >= QMAX:
                        stats["full"] += 1
                        break
                    q.append(int(neg.argmax()))
                    stepped += 1
                    stats["qhist"][len(q)] += 1
            live = [c for c in cells if not done[c]]
            if not live:
                break
            common = set(Q[live[0]])
            for c in live[1:]:
                common &= set(Q[c])
            common = sorted(common)
            if stepped == 0 and not common:
                stats["stuck"] += 1
                cnt = {}
                for c in live:
                    for q0 in Q[c]:
                        cnt[q0] = cnt.get(q0, 0) + 1
                qs = max(sorted(cnt), key=lambda k: cnt[k])
                g1 = [c for c in live if qs in Q[c]]
                g2 = [c for c in live if qs not in Q[c]]
                jobs.append((g2, S.copy(), mact.copy(), nmain))
                stats["snap"] += 1
                cells = g1
                common = set(Q[g1[0]])
                for c in g1[1:]:
                    common &= set(Q[c])
                common = sorted(common)
            else:
                cells = live
            stats["adv"].append(len(common))
            if common:
                stats["flush"] += 1
            for q0 in common:
                m = S[:, q0].copy()
                mu = m[q0]
                for c in cells:
                    wm[c] = wm[c] - (wm[c][q0] / mu) * m
                    wm[c][q0] = 0.0
                    Q[c].remove(q0)
                S -= np.outer(m, m / mu)
                S[q0, :] = 0.0
                S[:, q0] = 0.0
                mact[q0] = False
                nmain += 1
    stats["rounds"].append(rounds)
    return [(final[c], hand[c]) for c in range(C)]


stats = {"qhist": np.zeros(QMAX + 2, dtype=np.int64), "full": 0, "stuck": 0, "rounds": [], "adv": [], "qfinal": [0], "snap": 0, "flush": 0}
mism = 0
ncell = 0
for a in range(NTY):
    for b in range(NTX):
        rows = r0 + a * TY + np.arange(TY)
        cols = c0 + b * TX + np.arange(TX)
        rl = []
        for i in rows:
            for j in cols:
                X, Y = x0 + dx * j, y0 - dx * i
                r = np.ones(NN)
                r[:n] = np.sqrt((X - sx) ** 2 + (Y - sy) ** 2)
                rl.append(r)
        got = tile_walk(rl, stats)
        for r, g in zip(rl, got):
            ref = plain_walk(r)
            ncell += 1
            if ref != g:
                mism += 1
                print("  mismatch: plain", sorted(ref[0]), ref[1], " tile", sorted(g[0]), g[1])
qh = stats["qhist"]
print("tiles %dx%d of %dx%d cells at (%d,%d): RA=%d QMAX=%d" % (NTY, NTX, TY, TX, r0, c0, RA, QMAX))
print("cells %d, mismatches vs plain walk %d, full-Q stalls %d, splits %d, flushes %d" % (ncell, mism, stats["full"], stats["stuck"], stats["flush"]))
print("rounds per tile: mean %.1f max %d;  stations advanced per round: mean %.2f max %d"
      % (np.mean(stats["rounds"]), max(stats["rounds"]), np.mean(stats["adv"]), max(stats["adv"])))
print("|Q| after each step: " + "  ".join("%d:%.3f" % (i, qh[i] / max(1, qh.sum())) for i in range(1, QMAX + 1)))
print("mean |Q| %.2f;  largest Q left at the end of a tile: mean %.2f max %d"
      % ((qh * np.arange(len(qh))).sum() / max(1, qh.sum()), np.mean(stats["qfinal"]), max(stats["qfinal"])))
