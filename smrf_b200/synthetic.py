"""Seeded synthetic basins for parity tests and bench.py.

Everything here is NumPy on the host; it only *generates inputs* (DEM, station geometry, station
time series with dropouts, HRRR-like source lattices).  No interpolation happens here.
"""
import numpy as np


def make_dem(ny, nx, dx=30.0, x0=500000.0, y0=4800000.0, seed=20261019, noise=True):
    """Regular grid as topo files have it: x ascending, y DESCENDING; DEM float32 -> float64
    (smrf/data/load_topo.py:96-113 reads float32 NetCDF and casts)."""
    x = x0 + dx * np.arange(nx, dtype=np.float64)
    y = y0 - dx * np.arange(ny, dtype=np.float64)
    X = x[None, :]
    Y = y[:, None]
    dem = (1800.0
           + 700.0 * np.sin(2 * np.pi * X / 41000.0) * np.cos(2 * np.pi * Y / 57000.0)
           + 120.0 * np.sin(2 * np.pi * X / 3100.0 + 1.0) * np.sin(2 * np.pi * Y / 2300.0))
    if noise:
        rng = np.random.default_rng(seed)
        dem = dem + rng.normal(0.0, 8.0, size=dem.shape)
    dem = dem.astype(np.float32).astype(np.float64)
    return x, y, dem


def make_stations(S, x, y, dem, seed=7):
    """S stations uniform in the domain, shifted 0.37*dx off cell centres."""
    rng = np.random.default_rng(seed)
    nx, ny = len(x), len(y)
    dx = abs(x[1] - x[0]) if nx > 1 else 1.0
    dy = abs(y[1] - y[0]) if ny > 1 else dx
    j = rng.integers(0, nx, size=S)
    i = rng.integers(0, ny, size=S)
    mx = x[j] + 0.37 * dx
    my = y[i] + 0.37 * dy
    mz = dem[i, j] + rng.normal(0.0, 5.0, size=S)
    return mx, my, mz


def _ar1(T, S, rho, sigma, rng):
    e = rng.normal(0.0, sigma, size=(T, S))
    out = np.empty((T, S))
    out[0] = e[0]
    for t in range(1, T):
        out[t] = rho * out[t - 1] + e[t]
    return out


def make_series(kind, mz, T, seed=3):
    """Hourly station series [T, S] for 'air_temp' | 'vapor_pressure' | 'precip'."""
    rng = np.random.default_rng(seed)
    S = len(mz)
    t = np.arange(T, dtype=np.float64)[:, None]
    if kind == "air_temp":
        return (12.0 - 0.0065 * mz[None, :] + 8.0 * np.sin(2 * np.pi * t / 8760.0)
                + 5.0 * np.sin(2 * np.pi * t / 24.0) + _ar1(T, S, 0.9, 1.5, rng))
    if kind == "vapor_pressure":
        base = 900.0 * np.exp(-(mz[None, :] - 1000.0) / 2500.0)
        return base * rng.uniform(0.6, 1.0, size=(T, S))
    if kind == "precip":
        storm = rng.random(T) < 0.12
        amt = rng.gamma(0.8, 2.0, size=(T, S)) * (1.0 + 0.0008 * (mz[None, :] - 1800.0))
        amt = np.maximum(amt, 0.0)
        return np.where(storm[:, None], amt, 0.0)
    raise ValueError(kind)


def dropouts_iid(data, p=0.02, seed=11):
    """i.i.d. NaN with probability p per station-hour; never a fully-NaN row."""
    rng = np.random.default_rng(seed)
    out = data.copy()
    miss = rng.random(out.shape) < p
    allmiss = miss.all(axis=1)
    miss[allmiss, 0] = False
    out[miss] = np.nan
    return out


def dropouts_block(data, max_masks=64, mean_len=72.0, per_year=4.0, seed=13):
    """Block outages quantised so that the whole series has <= max_masks distinct NaN masks.

    Outage start/end times are snapped to a coarse time lattice of (max_masks-1) change points,
    so at most max_masks distinct station masks occur.  Returns (data_with_nan, n_distinct_masks).
    """
    rng = np.random.default_rng(seed)
    T, S = data.shape
    off = np.zeros((T, S), dtype=bool)
    rate = per_year / 8760.0
    for s in range(S):
        t = 0
        while True:
            t += int(np.ceil(rng.exponential(1.0 / rate)))
            if t >= T:
                break
            ln = int(np.ceil(rng.exponential(mean_len)))
            off[t:t + ln, s] = True
            t += ln
    nseg = max(1, max_masks)
    edges = np.linspace(0, T, nseg + 1).astype(int)
    q = np.zeros_like(off)
    for a, b in zip(edges[:-1], edges[1:]):
        if b > a:
            q[a:b] = off[a]          # one mask per segment: the state at the segment start
    q[q.all(axis=1), 0] = False
    out = data.copy()
    out[q] = np.nan
    nmask = len({row.tobytes() for row in q})
    return out, nmask


def make_source_lattice(x, y, dem, spacing=3000.0, jitter=150.0, seed=5):
    """HRRR-like source points: a regular lattice over (and slightly beyond) the DEM + jitter."""
    rng = np.random.default_rng(seed)
    xmin, xmax = min(x[0], x[-1]), max(x[0], x[-1])
    ymin, ymax = min(y[0], y[-1]), max(y[0], y[-1])
    cx, cy = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    nxs = int(np.ceil((xmax - xmin) / spacing)) + 3
    nys = int(np.ceil((ymax - ymin) / spacing)) + 3
    gx = cx + spacing * (np.arange(nxs) - (nxs - 1) / 2.0)
    gy = cy + spacing * (np.arange(nys) - (nys - 1) / 2.0)
    PX, PY = np.meshgrid(gx, gy)
    px = PX.ravel() + rng.uniform(-jitter, jitter, PX.size)
    py = PY.ravel() + rng.uniform(-jitter, jitter, PX.size)
    # elevation: nearest DEM cell (clamped to the domain) + noise, a stand-in for a 3 km smoothed DEM
    dx = abs(x[1] - x[0])
    dy = abs(y[1] - y[0])
    j = np.clip(np.rint((px - x[0]) / (x[1] - x[0])).astype(int), 0, len(x) - 1)
    i = np.clip(np.rint((py - y[0]) / (y[1] - y[0])).astype(int), 0, len(y) - 1)
    pz = dem[i, j] + rng.normal(0.0, 30.0, px.size)
    return px, py, pz


def make_source_fields(px, py, pz, T, seed=9):
    """Smooth synoptic gradient + lapse + noise on the source points, [T, P], no NaN."""
    rng = np.random.default_rng(seed)
    t = np.arange(T, dtype=np.float64)[:, None]
    gx = (px - px.mean()) / 50000.0
    gy = (py - py.mean()) / 50000.0
    return (10.0 + 3.0 * gx[None, :] * np.cos(2 * np.pi * t / 97.0) + 2.0 * gy[None, :] * np.sin(2 * np.pi * t / 61.0)
            - 0.0065 * pz[None, :] + 4.0 * np.sin(2 * np.pi * t / 24.0) + rng.normal(0.0, 0.4, size=(T, len(px))))
