"""Drop-in for smrf.spatial.grid.GRID running on libsmrfb (CUDA, sm_100a).

Same constructor and methods as the reference class (smrf/spatial/grid.py:23-231) as used by
smrf/distribute/image_data.py:189-195 and :245-255.  The reference re-triangulates and re-walks
every cell on every call (scipy.interpolate.griddata); here the Delaunay triangulation and the
simplex of every cell are built ONCE on the host with scipy -- the same qhull call the reference
makes -- and the per-timestep work (barycentric interpolation, retrend, clamp) is one CUDA kernel.

Covered: grid_method='linear', 'nearest' and 'cubic' with grid_local=False (detrendedInterpolationMask,
calculateInterpolation), and 'linear' / 'cubic' with grid_local=True (detrendedInterpolationLocal,
grid.py:115-177: the per-point local fits stay on the host -- P ~ 5000 points x k neighbours -- and the
three interpolations plus their combination are ONE kernel).

'cubic' (the reference's default grid_method; scipy's CloughTocher2DInterpolator): the vertex gradients are
estimated on the host by SciPy itself -- the estimator is an iteration whose stopping point decides the low
digits, so the same code has to run -- and the C1 Bezier patch, which is linear in the vertex values and
gradients, is evaluated on the device from nine per-cell weights formed once per handle.
"""
import os

import numpy as np

from .. import _lib


class GRID(_lib.HandleMixin):
    def __init__(self, config, mx, my, GridX, GridY, mz=None, GridZ=None, mask=None, metadata=None, device=None):
        from scipy.spatial import Delaunay

        self.config = config
        self.mx = _lib.as_f64(mx)
        self.my = _lib.as_f64(my)
        self.mz = None if mz is None else _lib.as_f64(mz)
        self.npoints = len(self.mx)
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.shape = tuple(np.shape(GridX))
        self.metadata = metadata
        self.pv = None
        self.grid_local = bool(config.get("grid_local", False))
        if self.grid_local:
            # neighbour table, grid.py:54-65: the k nearest points in latitude / longitude space, in the order of
            # pandas' Series.sort_values() (= numpy's default sort), as positions into the metadata frame
            k = int(config["grid_local_n"])
            lat = np.asarray(metadata.latitude.values, dtype=np.float64)
            lon = np.asarray(metadata.longitude.values, dtype=np.float64)
            self.nbr = np.empty((len(lat), k), dtype=np.int64)
            for i in range(len(lat)):
                d = np.sqrt((lat - lat[i]) ** 2 + (lon - lon[i]) ** 2)
                self.nbr[i] = np.argsort(d, kind="quicksort")[:k]
            self._elev = np.asarray(metadata.elevation.values, dtype=np.float64)
            # the local path triangulates the metadata's utm coordinates (grid.py:145-148)
            self.mx = _lib.as_f64(metadata.utm_x.values)
            self.my = _lib.as_f64(metadata.utm_y.values)

        # point mask, grid.py:80-94 (first minimum wins, like np.argmin)
        if config.get("grid_mask", False):
            assert mask.shape == self.shape
            m = np.asarray(mask).astype(bool)
            x = np.asarray(GridX)[0, :]
            y = np.asarray(GridY)[:, 0]
            xi = np.argmin(np.abs(x[None, :] - self.mx[:, None]), axis=1)
            yi = np.argmin(np.abs(y[None, :] - self.my[:, None]), axis=1)
            self.mask = m[yi, xi]
        else:
            self.mask = np.ones(self.npoints, dtype=bool)

        # triangulation + per-cell simplex, once (scipy.spatial.Delaunay default options, as griddata)
        self.tri = Delaunay(np.column_stack([self.mx, self.my]))
        kind, gx, gy = _lib.mesh_or_points(GridX, GridY)
        xi = np.column_stack([np.asarray(GridX, dtype=np.float64).ravel(), np.asarray(GridY, dtype=np.float64).ravel()])
        self.cell_simplex = np.ascontiguousarray(self.tri.find_simplex(xi).astype(np.int32))
        simplices = np.ascontiguousarray(self.tri.simplices.astype(np.int32))
        transform = np.ascontiguousarray(self.tri.transform, dtype=np.float64)
        pm = np.ascontiguousarray(self.mask.astype(np.uint8))
        dem = None if GridZ is None else _lib.as_f64(GridZ).reshape(-1)
        self.device = _lib.default_device() if device is None else int(device)
        self._h = _lib.handle_t()
        _lib.check(_lib.lib().smrfb_grid_create(
            self.npoints, _lib.dptr(self.mx), _lib.dptr(self.my), _lib.dptr(self.mz),
            self.shape[0], self.shape[1], kind, _lib.dptr(gx), _lib.dptr(gy), _lib.dptr(dem),
            simplices.shape[0], simplices.ctypes.data_as(_lib.c_i32_p), _lib.dptr(transform.reshape(-1)),
            self.cell_simplex.ctypes.data_as(_lib.c_i32_p), pm.ctypes.data_as(_lib.c_u8_p), self.device, self._h))

    def _method_code(self, grid_method):
        if grid_method == "linear":
            return 0
        if grid_method == "nearest":
            if not getattr(self, "_nearest_ready", False):
                # griddata(method='nearest') = NearestNDInterpolator: cKDTree(points).query(xi); same call, once
                from scipy.spatial import cKDTree
                xi = np.column_stack([np.asarray(self.GridX, dtype=np.float64).ravel(),
                                      np.asarray(self.GridY, dtype=np.float64).ravel()])
                idx = cKDTree(np.column_stack([self.mx, self.my])).query(xi)[1]
                self.cell_nearest = np.ascontiguousarray(idx.astype(np.int32))
                _lib.check(_lib.lib().smrfb_grid_set_nearest(self._h, self.cell_nearest.ctypes.data_as(_lib.c_i32_p)))
                self._nearest_ready = True
            return 1
        raise ValueError("unknown grid_method %r (the reference's griddata takes 'linear', 'nearest', 'cubic')" % (grid_method,))

    # ------------------------------------------------------------------ grid_method='cubic' (Clough-Tocher)
    def _cubic_geometry(self):
        """Per simplex the edge vectors e12, e23, e31 and the cross-edge coefficients g of SciPy's Clough-Tocher
        patch (the centroid of the neighbour across each edge in this simplex's barycentric coordinates; -1/2 on
        the hull), [nsimp, 9]."""
        tri = self.tri
        P, S, Tm = tri.points, tri.simplices, tri.transform
        e12 = P[S[:, 1]] - P[S[:, 0]]
        e23 = P[S[:, 2]] - P[S[:, 1]]
        e31 = P[S[:, 0]] - P[S[:, 2]]
        g = np.full((len(S), 3), -0.5)
        for k, (a, b) in enumerate(((2, 1), (0, 2), (1, 0))):
            nb = tri.neighbors[:, k]
            has = nb >= 0
            y = P[S[np.where(has, nb, 0)]].sum(axis=1) / 3.0
            d = y - Tm[:, 2, :]
            c0 = Tm[:, 0, 0] * d[:, 0] + Tm[:, 0, 1] * d[:, 1]
            c1 = Tm[:, 1, 0] * d[:, 0] + Tm[:, 1, 1] * d[:, 1]
            c = np.stack([c0, c1, (1.0 - c0) - c1], axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                gk = (2 * c[:, a] + c[:, b] - 1) / (2 - 3 * c[:, a] - 3 * c[:, b])
            g[:, k] = np.where(has, gk, -0.5)
        return np.ascontiguousarray(np.concatenate([e12, e23, e31, g], axis=1))

    def _ensure_cubic(self):
        if not getattr(self, "_cubic_ready", False):
            from scipy.interpolate import CloughTocher2DInterpolator
            geo = self._cubic_geometry()
            _lib.check(_lib.lib().smrfb_grid_set_cubic(self._h, _lib.dptr(geo.reshape(-1))))
            # scipy.spatial.Delaunay fills several attributes lazily (vertex_neighbor_vertices, transform, ...) and
            # not thread-safely: touch them once here, single-threaded, before _gradients fans out over threads
            # (measured: 8 threads on a fresh triangulation segfault, after this warm-up they are bit-identical)
            CloughTocher2DInterpolator(self.tri, np.zeros(self.npoints))
            self._cubic_ready = True

    def _gradients(self, values):
        """[T, P] point values -> [T, P, 2] vertex gradients, exactly what the reference's
        CloughTocher2DInterpolator(tri, values) estimates for itself (default tol / maxiter); columns are
        independent, so chunks of steps run on a few host threads (the Cython loop releases the GIL)."""
        from concurrent.futures import ThreadPoolExecutor
        from scipy.interpolate import CloughTocher2DInterpolator

        T = values.shape[0]
        y = np.ascontiguousarray(values.T)                            # [P, T]

        def run(sl):
            return CloughTocher2DInterpolator(self.tri, np.ascontiguousarray(y[:, sl])).grad       # [P, n, 2]
        nthr = min(8, os.cpu_count() or 1, max(1, T // 8))
        if nthr <= 1:
            g = run(slice(0, T))
        else:
            edges = np.linspace(0, T, nthr + 1).astype(int)
            with ThreadPoolExecutor(nthr) as ex:
                g = np.concatenate(list(ex.map(run, [slice(a, b) for a, b in zip(edges[:-1], edges[1:])])), axis=1)
        return np.ascontiguousarray(np.transpose(g, (1, 0, 2)))      # [T, P, 2]

    def _run_cubic(self, fields, nf, pv, vmin, vmax, out, out_dtype):
        T = fields.shape[0]
        f32 = self._select_dtype(out_dtype, out)
        out = _lib.out_array(T, self.shape, out, f32)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_grid_distribute_cubic(self._h, _lib.dptr(fields.reshape(-1)), T, nf, _lib.dptr(pv),
                                                          lo, hi, _lib.vptr(out)))
        return out

    def _distribute_block_cubic(self, data, detrend, flag, vmin, vmax, pv, out, out_dtype):
        """detrendedInterpolationMask / calculateInterpolation with grid_method='cubic' for a block of steps."""
        self._ensure_cubic()
        T = data.shape[0]
        pvs = None
        vals = data
        if detrend:
            if pv is not None:
                pvs = _lib.as_f64(pv).reshape(T, 2)
            else:      # grid.py:188-201: np.polyfit over the masked points, as the reference calls it
                pvs = np.empty((T, 2))
                mzm = self.mz[self.mask]
                for t in range(T):
                    p = np.polyfit(mzm, data[t][self.mask], 1)
                    if (flag == 1 and p[0] < 0) or (flag == -1 and p[0] > 0):
                        p = np.array([0.0, 0.0])
                    pvs[t] = p
            vals = data - (self.mz[None, :] * pvs[:, 0:1] + pvs[:, 1:2])
        grad = self._gradients(vals)
        fields = np.empty((T, 1, 3, self.npoints))
        fields[:, 0, 0] = vals
        fields[:, 0, 1] = grad[:, :, 0]
        fields[:, 0, 2] = grad[:, :, 1]
        self.pv_block = pvs
        return self._run_cubic(fields, 1, pvs, vmin, vmax, out, out_dtype)

    def distribute_block(self, data, detrend=False, flag=0, vmin=None, vmax=None, pv=None, out=None, grid_method="linear",
                         out_dtype=None):
        """[T, npoints] -> float64 [T, ny, nx] in one launch; 'linear' / 'cubic': NaN outside the convex hull."""
        data = _lib.as_f64(data)
        if data.ndim == 1:
            data = data[None, :]
        T, S = data.shape
        if S != self.npoints:
            raise ValueError("data has %d points, GRID was built for %d" % (S, self.npoints))
        if grid_method == "cubic":
            return self._distribute_block_cubic(data, detrend, flag, vmin, vmax, pv, out, out_dtype)
        method = self._method_code(grid_method)
        f32 = self._select_dtype(out_dtype, out)
        out = _lib.out_array(T, self.shape, out, f32)
        pv_in = None if pv is None else _lib.as_f64(pv).reshape(T, 2)
        pv_out = np.empty((T, 2), dtype=np.float64)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_grid_distribute(
            self._h, _lib.dptr(data), T, int(bool(detrend)), int(flag), method, lo, hi,
            _lib.dptr(pv_in), _lib.dptr(pv_out), _lib.vptr(out)))
        self.pv_block = pv_out
        return out

    # ------------------------------------------------------------------ grid_local (grid.py:115-177)
    def local_fit(self, data, flag=0, exact=True):
        """Per-point local trend of a block [T, npoints]: (residual, slope, intercept), each [T, npoints].

        exact=True calls np.polyfit per point and step like the reference's groupby/apply (grid.py:125-128), so the
        coefficients are bit-identical; exact=False is the closed-form least-squares line, vectorised over the
        block (agrees to ~1e-13 relative; use it for long blocks)."""
        data = _lib.as_f64(data)
        if data.ndim == 1:
            data = data[None, :]
        T, P = data.shape
        ze = self._elev[self.nbr]                                   # [P, k]
        slope = np.empty((T, P))
        icpt = np.empty((T, P))
        if exact:
            for t in range(T):
                zd = data[t][self.nbr]
                for i in range(P):
                    slope[t, i], icpt[t, i] = np.polyfit(ze[i], zd[i], 1)
        else:
            zm = ze.mean(axis=1)
            zc = ze - zm[:, None]
            szz = (zc * zc).sum(axis=1)
            dn = data[:, self.nbr]                                  # [T, P, k]
            dm = dn.mean(axis=2)
            slope = ((dn - dm[:, :, None]) * zc[None]).sum(axis=2) / szz[None]
            icpt = dm - slope * zm[None]
        if flag == 1:
            z = slope < 0
        elif flag == -1:
            z = slope > 0
        else:
            z = np.zeros_like(slope, dtype=bool)
        slope[z] = 0.0
        icpt[z] = 0.0
        first = self.nbr[:, 0]                                      # drop_duplicates(keep='first'), grid.py:131-133
        res = data[:, first] - (self._elev[first][None] * slope + icpt)
        return res, slope, icpt

    def distribute_block_local(self, data, flag=0, vmin=None, vmax=None, out=None, exact=False, grid_method="linear",
                               out_dtype=None):
        """[T, npoints] -> float64 [T, ny, nx]: detrendedInterpolationLocal for a block of steps, one launch."""
        res, slope, icpt = self.local_fit(data, flag, exact=exact)
        T = res.shape[0]
        if grid_method == "cubic":
            # three Clough-Tocher interpolations (utils.py:447), each with its own gradient estimate like the
            # reference's three CloughTocher2DInterpolator objects; the patch is linear in values and gradients, so
            # residual + intercept fold into one field and the kernel evaluates two: out = patch(0) + GridZ * patch(1)
            self._ensure_cubic()
            g = self._gradients(np.concatenate([res, icpt, slope], axis=0))          # [3T, P, 2]
            fields = np.empty((T, 2, 3, self.npoints))
            fields[:, 0, 0] = res + icpt
            fields[:, 0, 1] = g[:T, :, 0] + g[T:2 * T, :, 0]
            fields[:, 0, 2] = g[:T, :, 1] + g[T:2 * T, :, 1]
            fields[:, 1, 0] = slope
            fields[:, 1, 1] = g[2 * T:, :, 0]
            fields[:, 1, 2] = g[2 * T:, :, 1]
            return self._run_cubic(fields, 2, None, vmin, vmax, out, out_dtype)
        if grid_method != "linear":
            raise ValueError("grid_local takes grid_method 'linear' or 'cubic' (utils.py:430-449), got %r" % (grid_method,))
        f3 = np.ascontiguousarray(np.stack([res, slope, icpt], axis=1))      # [T, 3, P]
        f32 = self._select_dtype(out_dtype, out)
        out = _lib.out_array(T, self.shape, out, f32)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        # exact fits (np.polyfit per point, the per-step reference path) keep griddata's operation order bit for bit; the
        # vectorised closed-form fits of the block path differ from polyfit in the last digits anyway: folded, fused kernel
        _lib.check(_lib.lib().smrfb_grid_set_local_exact(self._h, 1 if exact else 0))
        _lib.check(_lib.lib().smrfb_grid_distribute_local(self._h, _lib.dptr(f3), T, lo, hi, _lib.vptr(out)))
        return out

    def detrendedInterpolationLocal(self, data, flag=0, grid_method="linear"):
        """grid.py:115-177.  ``data``: pandas Series indexed like the metadata, or an array in metadata order."""
        if hasattr(data, "loc") and self.metadata is not None:
            data = data.loc[self.metadata.index].values
        return self.distribute_block_local(np.asarray(data, dtype=np.float64)[None, :], flag=flag, exact=True,
                                           grid_method=grid_method)[0]

    def detrendedInterpolation(self, data, flag=0, grid_method="linear"):
        """grid.py:96-113 -> detrendedInterpolationLocal (grid_local) or detrendedInterpolationMask (grid.py:179-212).
        ``data`` may be a pandas Series."""
        if self.grid_local:
            return self.detrendedInterpolationLocal(data, flag, grid_method)
        v = self.distribute_block(np.asarray(data, dtype=np.float64)[None, :], detrend=True, flag=flag,
                                  grid_method=grid_method)[0]
        self.pv = self.pv_block[0].copy()
        return v

    def detrendedInterpolationMask(self, data, flag=0, grid_method="linear"):
        """grid.py:179-212."""
        v = self.distribute_block(np.asarray(data, dtype=np.float64)[None, :], detrend=True, flag=flag,
                                  grid_method=grid_method)[0]
        self.pv = self.pv_block[0].copy()
        return v

    def calculateInterpolation(self, data, grid_method="linear"):
        """grid.py:214-231."""
        return self.distribute_block(np.asarray(data, dtype=np.float64)[None, :], detrend=False, grid_method=grid_method)[0]
