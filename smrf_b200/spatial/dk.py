"""Drop-in for smrf.spatial.dk.dk.DK running on libsmrfb (CUDA, sm_100a).

Same constructor and methods as the reference class (smrf/spatial/dk/dk.py:20-168) as used by
smrf/distribute/image_data.py:182-185 and :242-243: ``DK(mx, my, mz, GridX, GridY, GridZ, config)``
with ``config['detrend_slope']`` (``config['dk_ncores']`` is accepted and ignored -- the kriging
solve runs on the GPU), ``calculate(data)`` for one timestep, plus ``distribute_block`` for a
whole [T, nsta] block.  Kriging weights are built on the device the first time a NaN mask is seen
and cached per mask (the reference caches exactly one mask, dk.py:78-85).

Differences, deliberate: ``weights`` (dense [ny, nx, nvalid], dk.py:133-137) is materialised
only on demand by ``calculateWeights()``; the distribute path keeps them sparse on the device.
"""
import numpy as np

from .. import _lib


class DK(_lib.HandleMixin):
    def __init__(self, mx, my, mz, GridX, GridY, GridZ, config, device=None):
        self.mx = _lib.as_f64(mx)
        self.my = _lib.as_f64(my)
        self.mz = _lib.as_f64(mz)
        self.nsta = len(self.mx)
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.shape = tuple(np.shape(GridX))
        self.ngrid = self.shape[0] * self.shape[1]
        self.config = config
        self.data = None
        self.nan_val = []
        self.weights = []
        self.pv = None
        self.residuals = None
        self.device = _lib.default_device() if device is None else int(device)
        kind, gx, gy = _lib.mesh_or_points(GridX, GridY)
        dem = _lib.as_f64(GridZ).reshape(-1)
        self._h = _lib.handle_t()
        _lib.check(_lib.lib().smrfb_dk_create(
            self.nsta, _lib.dptr(self.mx), _lib.dptr(self.my), _lib.dptr(self.mz),
            self.shape[0], self.shape[1], kind, _lib.dptr(gx), _lib.dptr(gy), _lib.dptr(dem),
            int(config.get("detrend_slope", 0) or 0), self.device, self._h))

    # ------------------------------------------------------------------ batched extension
    def distribute_block(self, data, vmin=None, vmax=None, pv=None, out=None, out_dtype=None):
        """[T, nsta] (NaN = missing) -> float64 [T, ny, nx]. Always detrends (image_data.py:242-243)."""
        data = _lib.as_f64(data)
        if data.ndim == 1:
            data = data[None, :]
        T, S = data.shape
        if S != self.nsta:
            raise ValueError("data has %d stations, DK was built for %d" % (S, self.nsta))
        f32 = self._select_dtype(out_dtype, out)
        out = _lib.out_array(T, self.shape, out, f32)
        pv_in = None if pv is None else _lib.as_f64(pv).reshape(T, 2)
        pv_out = np.empty((T, 2), dtype=np.float64)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_dk_distribute(self._h, _lib.dptr(data), T, lo, hi, _lib.dptr(pv_in),
                                                  _lib.dptr(pv_out), _lib.vptr(out)))
        self.pv_block = pv_out
        return out

    # ------------------------------------------------------------------ reference API
    def calculate(self, data):
        """dk.py:62-96."""
        data = np.asarray(data, dtype=np.float64)
        self.nan_val = np.isnan(data)
        v = self.distribute_block(data[None, :])[0]
        self.pv = self.pv_block[0].copy()
        ok = ~self.nan_val
        self.residuals = data[ok] - (self.mz[ok] * self.pv[0] + self.pv[1])
        return v

    def calculateWeights(self):
        """dk.py:98-137: dense kriging weights [ny, nx, nvalid] for ``self.nan_val`` (built on the GPU)."""
        nan_val = np.asarray(self.nan_val, dtype=bool)
        if nan_val.shape != (self.nsta,):
            nan_val = np.zeros(self.nsta, dtype=bool)
            self.nan_val = nan_val
        nvalid = int((~nan_val).sum())
        wg = np.zeros((self.ngrid, nvalid), dtype=np.float64)
        m = np.ascontiguousarray(nan_val.astype(np.uint8))
        _lib.check(_lib.lib().smrfb_dk_weights(self._h, m.ctypes.data_as(_lib.c_u8_p), _lib.dptr(wg)))
        self.weights = wg.reshape(self.shape + (nvalid,))
        return self.weights

    def detrendData(self, data):
        """dk.py:139-160 (host-side bookkeeping of an [nsta] vector; the kernels do this on the device)."""
        ok = ~np.asarray(self.nan_val, dtype=bool)
        pv = np.polyfit(self.mz[ok], np.asarray(data, dtype=np.float64)[ok], 1)
        flag = self.config.get("detrend_slope", 0)
        if (flag == 1 and pv[0] < 0) or (flag == -1 and pv[0] > 0):
            pv = np.array([0.0, 0.0])
        self.pv = pv
        self.residuals = np.asarray(data, dtype=np.float64)[ok] - (self.mz[ok] * pv[0] + pv[1])

    def retrendData(self, r):
        """dk.py:162-168."""
        return r + self.pv[0] * self.GridZ + self.pv[1]

    def stats(self):
        """(masks built, cells finished by the exact-solve continuation, largest final active set)."""
        import ctypes
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
        _lib.check(_lib.lib().smrfb_dk_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value
