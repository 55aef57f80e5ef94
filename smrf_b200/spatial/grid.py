"""Drop-in for smrf.spatial.grid.GRID (linear method) running on libsmrfb (CUDA, sm_100a).

Same constructor and methods as the reference class (smrf/spatial/grid.py:23-231) as used by
smrf/distribute/image_data.py:189-195 and :245-255.  The reference re-triangulates and re-walks
every cell on every call (scipy.interpolate.griddata); here the Delaunay triangulation and the
simplex of every cell are built ONCE on the host with scipy -- the same qhull call the reference
makes -- and the per-timestep work (barycentric interpolation, retrend, clamp) is one CUDA kernel.

Covered: grid_method='linear' and 'nearest' with grid_local=False (detrendedInterpolationMask,
calculateInterpolation), and grid_method='linear' with grid_local=True (detrendedInterpolationLocal,
grid.py:115-177: the per-point local fits stay on the host -- P ~ 5000 points x k neighbours -- and the
three interpolations plus their combination are ONE kernel).  'cubic' raises NotImplementedError
(SURVEY.md section 8f row N-b).
"""
import numpy as np

from .. import _lib


class GRID:
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

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().smrfb_destroy(h)
            except Exception:
                pass
            self._h = None

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
        raise NotImplementedError("grid_method=%r: 'linear' and 'nearest' are on the B200 path (SURVEY.md 8f N-b)" % grid_method)

    def distribute_block(self, data, detrend=False, flag=0, vmin=None, vmax=None, pv=None, out=None, grid_method="linear"):
        """[T, npoints] -> float64 [T, ny, nx] in one launch; 'linear': NaN outside the convex hull."""
        method = self._method_code(grid_method)
        data = _lib.as_f64(data)
        if data.ndim == 1:
            data = data[None, :]
        T, S = data.shape
        if S != self.npoints:
            raise ValueError("data has %d points, GRID was built for %d" % (S, self.npoints))
        if out is None:
            out = np.empty((T,) + self.shape, dtype=np.float64)
        pv_in = None if pv is None else _lib.as_f64(pv).reshape(T, 2)
        pv_out = np.empty((T, 2), dtype=np.float64)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_grid_distribute(
            self._h, _lib.dptr(data), T, int(bool(detrend)), int(flag), method, lo, hi,
            _lib.dptr(pv_in), _lib.dptr(pv_out), _lib.dptr(out)))
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

    def distribute_block_local(self, data, flag=0, vmin=None, vmax=None, out=None, exact=False):
        """[T, npoints] -> float64 [T, ny, nx]: detrendedInterpolationLocal for a block of steps, one launch."""
        res, slope, icpt = self.local_fit(data, flag, exact=exact)
        T = res.shape[0]
        f3 = np.ascontiguousarray(np.stack([res, slope, icpt], axis=1))      # [T, 3, P]
        if out is None:
            out = np.empty((T,) + self.shape, dtype=np.float64)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_grid_distribute_local(self._h, _lib.dptr(f3), T, lo, hi, _lib.dptr(out)))
        return out

    def detrendedInterpolationLocal(self, data, flag=0, grid_method="linear"):
        """grid.py:115-177.  ``data``: pandas Series indexed like the metadata, or an array in metadata order."""
        if grid_method != "linear":
            raise NotImplementedError("grid_local with grid_method=%r: only 'linear' is on the B200 path" % grid_method)
        if hasattr(data, "loc") and self.metadata is not None:
            data = data.loc[self.metadata.index].values
        return self.distribute_block_local(np.asarray(data, dtype=np.float64)[None, :], flag=flag, exact=True)[0]

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
