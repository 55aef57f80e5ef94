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
    """Reference GRID restricted to what the path covers: linear / nearest, mask and local detrend.

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
        """grid.py:115-177 with grid_method='linear' (utils.py:430-449: LinearNDInterpolator on the cached tri)."""
        from scipy.interpolate import LinearNDInterpolator
        from scipy.spatial import Delaunay

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
