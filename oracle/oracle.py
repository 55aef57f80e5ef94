"""CPU oracle for the SMRF spatial-distribution hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy (+ ``oracle/krige_oracle.c`` for the kriging solve), what the
reference computes for IDW, detrended kriging (DK), GRID-linear and the min/max clamp.  It is the
*checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing in ``smrf_b200/`` imports it, and the product
path fails loudly without its CUDA library rather than falling back to this.

Parity pinning: ``tests/test_oracle.py`` checks every function here against
``tests/golden/*.npz`` -- outputs of the reference's own ``idw.py`` / ``dk.py`` (+ its compiled C) /
``grid.py`` executed in the build container by ``tests/golden/make_golden.py`` -- and the kriging
restatement bit-for-bit against ``oracle/_ref/libkrige_ref.so`` (reference C compiled in place).

Reference citations are relative to ``/root/reference``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------------
# shared: linear elevation trend and clamp
# --------------------------------------------------------------------------------------------
def fit_trend(mz, data, flag):
    """Trend line through (mz, data) with the sign constraint.

    smrf/spatial/idw.py:119-129, smrf/spatial/dk/dk.py:146-155, smrf/spatial/grid.py:188-197:
    ``np.polyfit(..., 1)`` and, when the slope sign contradicts ``flag`` (+1 / -1), BOTH
    coefficients become 0.
    """
    pv = np.polyfit(np.asarray(mz, dtype=float), np.asarray(data, dtype=float), 1)
    if flag == 1 and pv[0] < 0:
        pv = np.array([0.0, 0.0])
    elif flag == -1 and pv[0] > 0:
        pv = np.array([0.0, 0.0])
    return pv


def set_min_max(data, min_val, max_val):
    """NaN-preserving clamp, in place (smrf/utils/utils.py:137-161)."""
    if max_val is None:
        max_val = np.inf
    if min_val is None:
        min_val = -np.inf
    keep_nan = np.isnan(data)
    data[data <= min_val] = min_val
    data[data >= max_val] = max_val
    data[keep_nan] = np.nan
    return data


# --------------------------------------------------------------------------------------------
# IDW  (smrf/spatial/idw.py)
# --------------------------------------------------------------------------------------------
class IDW:
    """Same constructor / methods as the reference class (idw.py:14-144)."""

    def __init__(self, mx, my, GridX, GridY, mz=None, GridZ=None, power=2, zeroVal=-1):
        self.mx, self.my, self.mz = mx, my, mz
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.power = power
        self.zeroVal = zeroVal
        self.npoints = len(mx)
        # idw.py:53-69 -- station-to-cell distances, station axis last
        dist = np.empty(GridX.shape + (self.npoints,))
        for s in range(self.npoints):
            dist[..., s] = np.sqrt((GridX - mx[s]) ** 2 + (GridY - my[s]) ** 2)
        # idw.py:69 -- the "zero fix" takes the min while the zeros are still there => no-op
        dist[np.where(dist == 0)] = np.min(dist)
        self.distance = dist
        # idw.py:71-77
        self.weights = 1.0 / (np.power(dist, power))

    def calculateIDW(self, data, local=False):
        # idw.py:82-94
        ok = ~np.isnan(data)
        w = self.weights[:, :, ok]
        d = data[ok]
        return np.nansum(w * d, 2) / np.sum(w, 2)

    def detrendData(self, data, flag=0, zeros=None):
        # idw.py:112-136 (fit on non-NaN stations; residual for every station)
        bad = np.isnan(data)
        self.pv = fit_trend(self.mz[~bad], data[~bad], flag)
        trend = self.mz * self.pv[0] + self.pv[1]
        if zeros is not None:
            data[zeros] = self.zeroVal
        self.dtrend = data - trend

    def retrendData(self, idw):
        # idw.py:138-144 (left to right)
        return idw + self.pv[0] * self.GridZ + self.pv[1]

    def detrendedIDW(self, data, flag=0, zeros=None, local=False):
        # idw.py:96-110
        self.detrendData(data, flag, zeros)
        v = self.retrendData(self.calculateIDW(self.dtrend, local))
        if zeros is not None:
            v[v < 0] = 0
        return v


# --------------------------------------------------------------------------------------------
# DK  (smrf/spatial/dk/dk.py + krige.c + lusolv.c)
# --------------------------------------------------------------------------------------------
_krige_lib = None
_krige_ref = None


def _load(path):
    lib = ctypes.CDLL(path)
    return lib


def krige_lib():
    """Our C restatement (built by oracle/build.py)."""
    global _krige_lib
    if _krige_lib is None:
        path = os.path.join(_HERE, "libkrige_oracle.so")
        if not os.path.exists(path):
            from . import build
            build.build_oracle()
        lib = _load(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.orc_krige_grid.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_int, dp, ip]
        lib.orc_krige_grid.restype = ctypes.c_int
        lib.orc_lusolv.argtypes = [ctypes.c_int, dp, dp]
        lib.orc_lusolv.restype = ctypes.c_int
        _krige_lib = lib
    return _krige_lib


def krige_ref_lib():
    """The reference's own C compiled in place (oracle/_ref), or None when absent."""
    global _krige_ref
    if _krige_ref is None:
        path = os.path.join(_HERE, "_ref", "libkrige_ref.so")
        if not os.path.exists(path):
            return None
        lib = _load(path)
        dp = ctypes.POINTER(ctypes.c_double)
        # dk_header.h:23
        lib.krige_grid.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_int, dp]
        lib.krige_grid.restype = None
        _krige_ref = lib
    return _krige_ref


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def call_grid(ad, dgrid, elevations, weights, nthreads=1, active=None, impl="oracle"):
    """``detrended_kriging.call_grid`` (detrended_kriging.pyx:32-67): weights filled in place.

    ``impl='oracle'`` runs krige_oracle.c, ``impl='ref'`` the compiled reference C.
    ``active`` (int32 [ngrid, nsta], oracle only) receives the final active-station flags.
    """
    ad = np.ascontiguousarray(ad, dtype=np.float64)
    dgrid = np.ascontiguousarray(dgrid, dtype=np.float64)
    elevations = np.ascontiguousarray(elevations, dtype=np.float64)
    assert weights.flags["C_CONTIGUOUS"] and weights.dtype == np.float64
    ngrid, nsta = dgrid.shape
    if impl == "ref":
        lib = krige_ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref/libkrige_ref.so not built")
        lib.krige_grid(nsta, ngrid, _dptr(ad), _dptr(dgrid), _dptr(elevations), int(nthreads), _dptr(weights))
        return None
    lib = krige_lib()
    ap = None
    if active is not None:
        assert active.dtype == np.int32 and active.flags["C_CONTIGUOUS"]
        ap = active.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    nbad = lib.orc_krige_grid(nsta, ngrid, _dptr(ad), _dptr(dgrid), _dptr(elevations), int(nthreads),
                              _dptr(weights), ap)
    if nbad:
        raise RuntimeError("singular kriging system in %d cells" % nbad)
    return None


def lusolv(a_aug):
    """Solve an n x (n+1) augmented system the way lusolv.c:26-48 does."""
    a = np.array(a_aug, dtype=np.float64, order="C")
    n = a.shape[0]
    x = np.zeros(n)
    rc = krige_lib().orc_lusolv(n, _dptr(a), _dptr(x))
    if rc:
        raise RuntimeError("singular")
    return x


def station_distances(mx, my):
    """dk.py:108-114 -- symmetric inter-station distance matrix, zero diagonal."""
    n = len(mx)
    ad = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            ad[i, j] = ad[j, i] = np.sqrt((mx[i] - mx[j]) ** 2 + (my[i] - my[j]) ** 2)
    return ad


class DK:
    """Same constructor / methods as the reference class (dk.py:20-168)."""

    def __init__(self, mx, my, mz, GridX, GridY, GridZ, config, impl="oracle"):
        self.mx, self.my, self.mz = mx, my, mz
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.nsta = len(mx)
        self.ngrid = GridX.shape[0] * GridX.shape[1]
        self.config = config
        self.nan_val = []
        self.weights = []
        self.impl = impl

    def calculateWeights(self):
        # dk.py:98-137
        ok = ~self.nan_val
        mx, my, mz = self.mx[ok], self.my[ok], self.mz[ok]
        nsta = int(ok.sum())
        self.ad = station_distances(mx, my)
        Xa, Ya = self.GridX.ravel(), self.GridY.ravel()
        dgrid = np.zeros((self.ngrid, nsta))
        for i in range(nsta):
            dgrid[:, i] = np.sqrt((Xa - mx[i]) ** 2 + (Ya - my[i]) ** 2)
        self.dgrid = dgrid
        wg = np.zeros_like(dgrid)
        call_grid(self.ad, dgrid, mz.astype(np.double), wg, self.config["dk_ncores"], impl=self.impl)
        self.weights = wg.reshape(self.GridX.shape + (nsta,))

    def detrendData(self, data):
        # dk.py:139-160 (residuals only for valid stations; flag from config['detrend_slope'])
        ok = ~self.nan_val
        self.pv = fit_trend(self.mz[ok], data[ok], self.config["detrend_slope"])
        self.residuals = data[ok] - (self.mz[ok] * self.pv[0] + self.pv[1])

    def retrendData(self, r):
        # dk.py:162-168
        return r + self.pv[0] * self.GridZ + self.pv[1]

    def calculate(self, data):
        # dk.py:62-96 (weights cached for one NaN mask)
        nan_val = np.isnan(np.asarray(data, dtype=float))
        if not np.array_equal(nan_val, self.nan_val):
            self.nan_val = nan_val
            self.calculateWeights()
        self.detrendData(data)
        r = np.nansum(self.weights * self.residuals, 2)
        return self.retrendData(r)


# --------------------------------------------------------------------------------------------
# GRID, linear  (smrf/spatial/grid.py + scipy.interpolate.griddata)
# --------------------------------------------------------------------------------------------
# The arithmetic of this method lives in SciPy (third-party; reference pin scipy>=1.4,<1.7 in
# requirements.txt, this image has 1.18): griddata(method='linear') = LinearNDInterpolator over
# scipy.spatial.Delaunay.  ``GRID`` below calls public griddata exactly where the reference does
# (grid.py:204-207, 226-229); ``linear_plan``/``linear_apply`` restate it in decomposed form --
# the form the CUDA path uses -- and tests pin the two against each other bit-for-bit.
def point_mask(mx, my, GridX, GridY, mask):
    """grid.py:80-94 -- mask value of the cell nearest each source point (first minimum on ties)."""
    x = GridX[0, :]
    y = GridY[:, 0]
    out = np.zeros(len(mx), dtype=bool)
    m = np.asarray(mask).astype(bool)
    for i in range(len(mx)):
        xi = np.argmin(np.abs(x - mx[i]))
        yi = np.argmin(np.abs(y - my[i]))
        out[i] = m[yi, xi]
    return out


def linear_plan(mx, my, GridX, GridY):
    """Delaunay + find_simplex once: per-cell simplex id (-1 outside the hull), vertex ids, transform."""
    from scipy.spatial import Delaunay

    tri = Delaunay(np.column_stack([np.asarray(mx, float), np.asarray(my, float)]))
    xi = np.column_stack([GridX.ravel(), GridY.ravel()])
    simplex = tri.find_simplex(xi).astype(np.int32)
    return {
        "simplex": simplex,                                    # int32 [ncell]
        "simplices": tri.simplices.astype(np.int32),           # int32 [nsimp, 3]
        "transform": np.ascontiguousarray(tri.transform),      # f64 [nsimp, 3, 2]
    }


def linear_apply(plan, values, GridX, GridY):
    """out = ((0 + c0*v_a) + c1*v_b) + c2*v_c with c from the simplex transform; NaN outside the hull.

    Evaluated with separate multiplies and adds (no FMA), the order that reproduced
    griddata(method='linear') bit-for-bit.
    """
    s = plan["simplex"]
    inside = s >= 0
    si = np.where(inside, s, 0)
    T = plan["transform"][si]
    x = GridX.ravel()
    y = GridY.ravel()
    dx = x - T[:, 2, 0]
    dy = y - T[:, 2, 1]
    c0 = T[:, 0, 0] * dx + T[:, 0, 1] * dy
    c1 = T[:, 1, 0] * dx + T[:, 1, 1] * dy
    c2 = (1.0 - c0) - c1
    v = np.asarray(values, dtype=float)[plan["simplices"][si]]
    out = ((0.0 + c0 * v[:, 0]) + c1 * v[:, 1]) + c2 * v[:, 2]
    out[~inside] = np.nan
    return out.reshape(GridX.shape)


# --------------------------------------------------------------------------------------------
# GRID, cubic  (grid_method='cubic': scipy.interpolate.CloughTocher2DInterpolator)
# --------------------------------------------------------------------------------------------
# Third-party algorithm (SciPy, _interpnd.pyx: `_clough_tocher_2d_single`, Alfeld 1984 / Farin 1986 with SciPy's
# affine-invariant choice of the cross-edge directions), reached from the reference at grid.py:204-207 / :226-229
# (griddata(method='cubic')) and utils.py:447 (CloughTocher2DInterpolator(tri, values)).  ``GRID`` calls SciPy
# where the reference does; ``cubic_plan`` / ``cubic_apply`` restate the evaluation in the decomposed form the
# CUDA path uses: the Bezier patch is LINEAR in the nine data of a simplex (three vertex values f_v and three vertex
# gradients), so per cell nine weights are formed once and a timestep costs nine multiply-adds.  The vertex gradients
# themselves come from SciPy's own estimator (estimate_gradients_2d_global, an iteration whose stopping point
# decides the low digits): both the oracle and the product call it, as the reference's CloughTocher2DInterpolator does.
def estimate_gradients(tri, values, tol=1e-6, maxiter=400):
    """[P] or [P, ncol] -> [P, ncol, 2]: what CloughTocher2DInterpolator(tri, values).grad holds."""
    from scipy.interpolate import _interpnd

    v = np.asarray(values, dtype=float)
    if v.ndim == 1:
        v = v[:, None]
    return _interpnd.estimate_gradients_2d_global(tri, np.ascontiguousarray(v), maxiter=maxiter, tol=tol)


def cubic_geometry(tri):
    """Per simplex: the edge vectors e12, e23, e31 and the three cross-edge coefficients g (centroid of the
    neighbour across each edge in this simplex's barycentric coordinates; -1/2 where there is no neighbour)."""
    P = tri.points
    S = tri.simplices
    e12 = P[S[:, 1]] - P[S[:, 0]]
    e23 = P[S[:, 2]] - P[S[:, 1]]
    e31 = P[S[:, 0]] - P[S[:, 2]]
    g = np.full((len(S), 3), -0.5)
    Tm = tri.transform
    for k in range(3):
        nb = tri.neighbors[:, k]
        has = nb >= 0
        y = P[S[np.where(has, nb, 0)]].sum(axis=1) / 3.0           # neighbour centroid
        d = y - Tm[:, 2, :]
        c0 = Tm[:, 0, 0] * d[:, 0] + Tm[:, 0, 1] * d[:, 1]
        c1 = Tm[:, 1, 0] * d[:, 0] + Tm[:, 1, 1] * d[:, 1]
        c = np.stack([c0, c1, (1.0 - c0) - c1], axis=1)
        a, b = [(2, 1), (0, 2), (1, 0)][k]
        gk = (2 * c[:, a] + c[:, b] - 1) / (2 - 3 * c[:, a] - 3 * c[:, b])
        g[:, k] = np.where(has, gk, -0.5)
    return np.concatenate([e12, e23, e31, g], axis=1)               # [nsimp, 9]


def _ct_patch(b, geo, f, df):
    """The Clough-Tocher patch value for barycentrics b [n,3], geometry geo [n,9], vertex values f [n,3] and vertex
    gradients df [n,3,2] -- SciPy's `_clough_tocher_2d_single`, vectorised over n."""
    e12x, e12y, e23x, e23y, e31x, e31y, g0, g1, g2 = [geo[:, i] for i in range(9)]
    f1, f2, f3 = f[:, 0], f[:, 1], f[:, 2]
    df12 = +(df[:, 0, 0] * e12x + df[:, 0, 1] * e12y)
    df21 = -(df[:, 1, 0] * e12x + df[:, 1, 1] * e12y)
    df23 = +(df[:, 1, 0] * e23x + df[:, 1, 1] * e23y)
    df32 = -(df[:, 2, 0] * e23x + df[:, 2, 1] * e23y)
    df31 = +(df[:, 2, 0] * e31x + df[:, 2, 1] * e31y)
    df13 = -(df[:, 0, 0] * e31x + df[:, 0, 1] * e31y)
    c3000 = f1
    c2100 = (df12 + 3 * c3000) / 3
    c2010 = (df13 + 3 * c3000) / 3
    c0300 = f2
    c1200 = (df21 + 3 * c0300) / 3
    c0210 = (df23 + 3 * c0300) / 3
    c0030 = f3
    c1020 = (df31 + 3 * c0030) / 3
    c0120 = (df32 + 3 * c0030) / 3
    c2001 = (c2100 + c2010 + c3000) / 3
    c0201 = (c1200 + c0300 + c0210) / 3
    c0021 = (c1020 + c0120 + c0030) / 3
    c0111 = (g0 * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) + (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2
    c1011 = (g1 * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) + (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2
    c1101 = (g2 * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) + (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2
    c1002 = (c1101 + c1011 + c2001) / 3
    c0102 = (c1101 + c0111 + c0201) / 3
    c0012 = (c1011 + c0111 + c0021) / 3
    c0003 = (c1002 + c0102 + c0012) / 3
    minval = b.min(axis=1)
    b1, b2, b3 = b[:, 0] - minval, b[:, 1] - minval, b[:, 2] - minval
    b4 = 3 * minval
    return (b1 ** 3 * c3000 + 3 * b1 ** 2 * b2 * c2100 + 3 * b1 ** 2 * b3 * c2010 + 3 * b1 ** 2 * b4 * c2001
            + 3 * b1 * b2 ** 2 * c1200 + 6 * b1 * b2 * b4 * c1101 + 3 * b1 * b3 ** 2 * c1020 + 6 * b1 * b3 * b4 * c1011
            + 3 * b1 * b4 ** 2 * c1002 + b2 ** 3 * c0300 + 3 * b2 ** 2 * b3 * c0210 + 3 * b2 ** 2 * b4 * c0201
            + 3 * b2 * b3 ** 2 * c0120 + 6 * b2 * b3 * b4 * c0111 + 3 * b2 * b4 ** 2 * c0102 + b3 ** 3 * c0030
            + 3 * b3 ** 2 * b4 * c0021 + 3 * b3 * b4 ** 2 * c0012 + b4 ** 3 * c0003)


def cubic_plan(mx, my, GridX, GridY, tri=None):
    """Delaunay + find_simplex once, then per cell the nine Clough-Tocher weights: out = sum_v (wf[v] f_v +
    wx[v] df_v/dx + wy[v] df_v/dy) over the simplex's three vertices; NaN outside the hull."""
    from scipy.spatial import Delaunay

    if tri is None:
        tri = Delaunay(np.column_stack([np.asarray(mx, float), np.asarray(my, float)]))
    xi = np.column_stack([GridX.ravel(), GridY.ravel()])
    simplex = tri.find_simplex(xi).astype(np.int32)
    inside = simplex >= 0
    si = np.where(inside, simplex, 0)
    Tm = tri.transform[si]
    dx = xi[:, 0] - Tm[:, 2, 0]
    dy = xi[:, 1] - Tm[:, 2, 1]
    c0 = Tm[:, 0, 0] * dx + Tm[:, 0, 1] * dy
    c1 = Tm[:, 1, 0] * dx + Tm[:, 1, 1] * dy
    b = np.stack([c0, c1, (1.0 - c0) - c1], axis=1)
    geo = cubic_geometry(tri)[si]
    n = len(si)
    w = np.empty((n, 9))
    for i in range(9):                                   # the patch is linear: evaluate it on the nine unit data
        f = np.zeros((n, 3))
        df = np.zeros((n, 3, 2))
        if i < 3:
            f[:, i] = 1.0
        else:
            df[:, (i - 3) // 2, (i - 3) % 2] = 1.0
        w[:, i] = _ct_patch(b, geo, f, df)
    return {"tri": tri, "simplex": simplex, "simplices": tri.simplices.astype(np.int32), "weights": w,
            "bary": b, "geo": geo}


def cubic_apply(plan, values, grad, shape):
    """values [P], grad [P, 2] -> [ny, nx]."""
    s = plan["simplex"]
    inside = s >= 0
    vt = plan["simplices"][np.where(inside, s, 0)]
    w = plan["weights"]
    v = np.asarray(values, dtype=float)
    g = np.asarray(grad, dtype=float)
    out = np.zeros(len(s))
    for k in range(3):
        out = out + w[:, k] * v[vt[:, k]]
    for k in range(3):
        out = out + w[:, 3 + 2 * k] * g[vt[:, k], 0] + w[:, 4 + 2 * k] * g[vt[:, k], 1]
    out[~inside] = np.nan
    return out.reshape(shape)


def local_neighbours(lat, lon, k):
    """grid.py:54-65: for every point the k nearest points in latitude / longitude space (itself first), in the
    order of pandas' Series.sort_values() = numpy's default (quick)sort of the distances."""
    lat = np.asarray(lat, dtype=float)
    lon = np.asarray(lon, dtype=float)
    nbr = np.empty((len(lat), k), dtype=np.int32)
    for i in range(len(lat)):
        d = np.sqrt((lat - lat[i]) ** 2 + (lon - lon[i]) ** 2)
        nbr[i] = np.argsort(d, kind="quicksort")[:k]
    return nbr


def local_fit(data, elevation, nbr, flag):
    """grid.py:125-141 and :164-166 for one timestep: per point a line through its k neighbours (np.polyfit on the
    neighbours in table order), the slope constraint, and the residual of the point's FIRST neighbour (the frame is
    de-duplicated with keep='first').  Returns dtrend, slope, intercept, each [P]."""
    data = np.asarray(data, dtype=float)
    elevation = np.asarray(elevation, dtype=float)
    P = nbr.shape[0]
    slope = np.empty(P)
    intercept = np.empty(P)
    for i in range(P):
        slope[i], intercept[i] = np.polyfit(elevation[nbr[i]], data[nbr[i]], 1)
    if flag == 1:
        z = slope < 0
    elif flag == -1:
        z = slope > 0
    else:
        z = np.zeros(P, dtype=bool)
    slope[z] = 0.0
    intercept[z] = 0.0
    first = nbr[:, 0]
    el_trend = elevation[first] * slope + intercept
    return data[first] - el_trend, slope, intercept


class GRID:
    """Reference GRID restricted to what the path covers: linear / nearest / cubic, mask and local detrend.

    grid.py:23-94 (ctor / neighbour table / point mask), :115-177 (detrendedInterpolationLocal), :179-212
    (detrendedInterpolationMask), :214-231.
    """

    def __init__(self, config, mx, my, GridX, GridY, mz=None, GridZ=None, mask=None, metadata=None):
        self.config = config
        self.mx, self.my, self.mz = mx, my, mz
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.npoints = len(mx)
        self.metadata = metadata
        if config.get("grid_local"):
            self.nbr = local_neighbours(metadata.latitude.values, metadata.longitude.values, config["grid_local_n"])
            self.tri = None
        if config["grid_mask"]:
            assert mask.shape == GridX.shape
            self.mask = point_mask(mx, my, GridX, GridY, mask)
        else:
            self.mask = np.ones_like(self.mx, dtype=bool)

    def detrendedInterpolation(self, data, flag=0, grid_method="linear"):
        from scipy.interpolate import griddata

        data = np.asarray(data, dtype=float)
        if self.config.get("grid_local"):
            return self.detrendedInterpolationLocal(data, flag, grid_method)
        self.pv = fit_trend(np.asarray(self.mz)[self.mask].astype(float), data[self.mask], flag)
        dtrend = data - (self.mz * self.pv[0] + self.pv[1])
        idtrend = griddata((self.mx, self.my), dtrend, (self.GridX, self.GridY), method=grid_method)
        return idtrend + self.pv[0] * self.GridZ + self.pv[1]

    def detrendedInterpolationLocal(self, data, flag=0, grid_method="linear"):
        """grid.py:115-177 (utils.py:430-449: LinearNDInterpolator / CloughTocher2DInterpolator on the cached tri)."""
        from scipy.interpolate import CloughTocher2DInterpolator, LinearNDInterpolator
        from scipy.spatial import Delaunay

        if grid_method == "cubic":
            LinearNDInterpolator = CloughTocher2DInterpolator          # noqa: F811  (utils.py:447)

        md = self.metadata
        dtrend, slope, intercept = local_fit(data, md.elevation.values, self.nbr, flag)
        if self.tri is None:
            self.tri = Delaunay(np.column_stack([md.utm_x.values, md.utm_y.values]))
        pts = (self.GridX, self.GridY)
        grid_slope = LinearNDInterpolator(self.tri, slope)(pts)
        grid_intercept = LinearNDInterpolator(self.tri, intercept)(pts)
        idtrend = LinearNDInterpolator(self.tri, dtrend)(pts)
        return idtrend + grid_slope * self.GridZ + grid_intercept

    def calculateInterpolation(self, data, grid_method="linear"):
        from scipy.interpolate import griddata

        return griddata((self.mx, self.my), np.asarray(data, dtype=float), (self.GridX, self.GridY),
                        method=grid_method)


# --------------------------------------------------------------------------------------------
# comparison metric
# --------------------------------------------------------------------------------------------
def parity_error(got, ref, scale):
    """max |got-ref| / max(|ref|, scale) over non-NaN cells; raises if NaN patterns differ."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    if not np.array_equal(np.isnan(got), np.isnan(ref)):
        raise AssertionError("NaN pattern differs: %d vs %d NaNs" % (np.isnan(got).sum(), np.isnan(ref).sum()))
    ok = ~np.isnan(ref)
    if not ok.any():
        return 0.0
    den = np.maximum(np.abs(ref[ok]), scale)
    return float(np.max(np.abs(got[ok] - ref[ok]) / den))


# --------------------------------------------------------------------------------------------
# consumers of a distributed grid (SURVEY.md 8f N-a) and WindNinja-style weights (8a G5)
# --------------------------------------------------------------------------------------------
def dew_point(vp, tol=0.01, impl="oracle"):
    """Dew point (degC) of vapor pressure (Pa): impl='oracle' -> oracle/envphys_oracle.c (restatement of dewpt.c:14-229),
    impl='ref' -> the reference's own dewpt.c + topotherm.c compiled in place (oracle/_ref/libenvphys_ref.so)."""
    vp = np.ascontiguousarray(vp, dtype=np.float64)
    out = np.empty_like(vp)
    if impl == "ref":
        path = os.path.join(_HERE, "_ref", "libenvphys_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libenvphys_ref.so is missing (python oracle/build.py in the build container)")
        L = _load(path)
        L.dewpt.restype = None
        L.dewpt.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        L.dewpt(int(vp.size), _dptr(vp.reshape(-1)), 1, float(tol), _dptr(out.reshape(-1)))     # dewpt.c:14-20
        return out
    path = os.path.join(_HERE, "libenvphys_oracle.so")
    if not os.path.exists(path):
        from . import build
        build.build_oracle()
    L = _load(path)
    L.oracle_dew_point.restype = None
    L.oracle_dew_point.argtypes = [ctypes.c_long, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    L.oracle_dew_point(int(vp.size), _dptr(vp.reshape(-1)), float(tol), _dptr(out.reshape(-1)))
    return out


def wet_bulb(ta, td, z, tol=0.01, impl="oracle"):
    """Wet / ice bulb temperature (degC) from air temperature, dew point (degC) and elevation (m): impl='oracle' ->
    oracle/envphys_oracle.c (restatement of iwbt.c:14-134), impl='ref' -> the reference's own iwbt.c compiled in place.  Inputs
    for which the reference would exit(-1) (below 0 K, no convergence) must not be given to impl='ref'."""
    ta, td, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (ta, td, z))
    out = np.empty_like(ta)
    dp = ctypes.POINTER(ctypes.c_double)
    if impl == "ref":
        path = os.path.join(_HERE, "_ref", "libenvphys_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libenvphys_ref.so is missing (python oracle/build.py in the build container)")
        L = _load(path)
        L.iwbt.restype = None
        L.iwbt.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_double, dp]
        L.iwbt(int(ta.size), _dptr(ta.reshape(-1)), _dptr(td.reshape(-1)), _dptr(z.reshape(-1)), 1, float(tol), _dptr(out.reshape(-1)))   # iwbt.c:84-91
        return out
    path = os.path.join(_HERE, "libenvphys_oracle.so")
    from . import build
    build.build_oracle()
    L = _load(path)
    L.oracle_wet_bulb.restype = ctypes.c_int
    L.oracle_wet_bulb.argtypes = [ctypes.c_long, dp, dp, dp, ctypes.c_double, dp]
    bad = L.oracle_wet_bulb(int(ta.size), _dptr(ta.reshape(-1)), _dptr(td.reshape(-1)), _dptr(z.reshape(-1)), float(tol), _dptr(out.reshape(-1)))
    if bad:
        raise RuntimeError("wet bulb: the reference would exit here (flags %d: 1 = no convergence in 10 steps, 2 = below 0 K)" % bad)
    return out


def dew_point_adjust(dpt, ta):
    """vapor_pressure.py:143-147: the dew point may not exceed the air temperature."""
    dpt = dpt.copy()
    ind = dpt >= ta
    if np.sum(ind) > 0:
        dpt[ind] = ta[ind] - 0.2
    return dpt


def snow_phase_and_density(temperature, precipitation, nasde_model):
    """envphys/snow.py:37-75 Snow.phase_and_density -> (snow_density, percent_snow) of the three new-accumulated-snow-density
    models: 'susong1999' (nasde_model/susong_1999.py:29-116, a 7-row table of the precipitation temperature),
    'piecewise_susong1999' (piecewise_suosong_1999.py:42-87 with utils.py:12-71) and 'marks2017' (marks_2017.py:73-207:
    compaction by overburden and metamorphism where there is snow water equivalent).  The inputs are copied, as the
    reference's np.array() does (the models clip the temperature in place)."""
    pp = np.array(precipitation, dtype=float)
    t = np.array(temperature, dtype=float)
    if nasde_model == "susong1999":
        ps = np.zeros(pp.shape)
        sd = np.zeros(pp.shape)
        if np.sum(pp) == 0:                                         # susong_1999.py:96-97
            return sd, ps
        rows = [(-np.inf, -5, 1, 75), (-5, -3, 1, 100), (-3, -1.5, 1, 150), (-1.5, -0.5, 1, 175), (-0.5, 0.0, 0.75, 200),
                (0.0, 0.5, 0.25, 250), (0.5, np.inf, 0, 0)]          # susong_1999.py:29-72; 1e309 is inf
        for lo, hi, snow, den in rows:
            ind = (t >= lo) & (t < hi)
            ps[ind] = snow
            sd[ind] = den
        ps[pp == 0] = 0                                             # susong_1999.py:111-112
        sd[pp == 0] = 0
        return sd, ps
    if nasde_model not in ("piecewise_susong1999", "marks2017"):
        raise ValueError("%s is not an implemented new accumulated snow density (NASDE) model" % nasde_model)

    def piecewise(tpp):                                             # check_temperature + calc_percent_snow + density
        tpp[tpp < -10.0] = -10.0                                    # utils.py:67-70 (t_min = -10, t_max = 0)
        tsnow = tpp.copy()
        tsnow[tpp > 0.0] = 0.0
        pcs = np.zeros(tpp.shape)                                   # utils.py:25-37
        pcs[tpp <= -0.5] = 1.0
        ind = (tpp > -0.5) & (tpp <= 0.0)
        pcs[ind] = (-tpp[ind] / 0.5) * 0.25 + 0.75
        ind = (tpp > 0.0) & (tpp <= 0.0 + 1.0)
        pcs[ind] = (-tpp[ind] / (0.0 + 1.0)) * 0.75 + 0.75
        t_range = 0.0 - -10.0                                       # piecewise_suosong_1999.py:74-83
        ex = 1.0 + (((t_range + (tsnow - 0.0)) / t_range) * 0.75)
        ex[ex > 1.75] = 1.75
        rho_ns = 50.0 + (1.7 * ((tpp - 0.0) + 15.0) ** ex)
        return pcs, rho_ns, tsnow

    if nasde_model == "piecewise_susong1999":
        pcs, rho_ns, _ = piecewise(t)
        return rho_ns, pcs
    # marks2017: rho_ns, d_rho_m, d_rho_c, zs and rho_s are ONE array in the reference (the chained assignment of
    # marks_2017.py:113-115), so the last write -- rho_s -- is what 'rho_s' returns: 0 where there is no snow water equivalent
    rho_s = np.zeros(t.shape)
    pcs = np.zeros(t.shape)
    if np.any(pp > 0):                                              # marks_2017.py:118
        pcs, rho_orig, tsnow = piecewise(t)
        swe = pp * pcs
        i = swe > 0.0
        if np.any(i):
            s_tsnow = tsnow[i]
            s_swe = swe[i]
            s_rho_ns = rho_orig[i] / 1000.0
            d_rho_c = (0.026 * np.exp(-0.08 * (0.0 - s_tsnow)) * s_swe * np.exp(-21.0 * s_rho_ns))      # :149-150
            ind = (s_rho_ns * 1000.0) >= 100.0
            c11 = np.ones(s_rho_ns.shape)
            c11[ind] = np.exp(-0.046 * (s_rho_ns[ind] * 1000.0 - 100.0))                                 # :155
            d_rho_m = 0.01 * c11 * np.exp(-0.04 * (0.0 - s_tsnow))                                       # :157-159
            rho_s[i] = s_rho_ns + ((d_rho_c + d_rho_m) * s_rho_ns)                                       # :162
    return rho_s * 1000.0, pcs


def wind_direction(u, v):
    """wind.py:175-178."""
    az = np.arctan2(u, v) * 180 / np.pi
    az[az < 0] = az[az < 0] + 360
    return az


def interp_weights(xy, uv, d=2):
    """utils.py:377-406."""
    from scipy.spatial import Delaunay
    tri = Delaunay(xy)
    simplex = tri.find_simplex(uv)
    vertices = np.take(tri.simplices, simplex, axis=0)
    temp = np.take(tri.transform, simplex, axis=0)
    delta = uv - temp[:, d]
    bary = np.einsum('njk,nk->nj', temp[:, :d, :], delta)
    return vertices, np.hstack((bary, 1 - bary.sum(axis=1, keepdims=True)))


def grid_interpolate(values, vtx, wts, shp, fill_value=np.nan):
    """utils.py:409-427."""
    ret = np.einsum('nj,nj->n', np.take(values, vtx), wts)
    ret[np.any(wts < 0, axis=1)] = fill_value
    return ret.reshape(shp[0], shp[1])
