"""Drop-in for smrf.spatial.idw.IDW running on libsmrfb (CUDA, sm_100a).

Same constructor and methods as the reference class (smrf/spatial/idw.py:14-144) as used by
smrf/distribute/image_data.py:176-178 and :233-240, plus ``distribute_block`` which distributes
a whole [T, nsta] block of timesteps in one fused launch (the reason this exists).
Differences, all deliberate:
  * the [ny, nx, nsta] ``distance`` / ``weights`` cubes are never materialised (they are
    regenerated per 64-cell tile in shared memory), so those two attributes do not exist;
  * the caller's ``data`` is never mutated.
"""
import numpy as np

from .. import _lib


class IDW(_lib.HandleMixin):
    def __init__(self, mx, my, GridX, GridY, mz=None, GridZ=None, power=2, zeroVal=-1, device=None):
        self.mx = _lib.as_f64(mx)
        self.my = _lib.as_f64(my)
        self.mz = None if mz is None else _lib.as_f64(mz)
        self.npoints = len(self.mx)
        self.GridX, self.GridY, self.GridZ = GridX, GridY, GridZ
        self.shape = tuple(np.shape(GridX))
        self.power = power
        self.zeroVal = zeroVal
        self.data = None
        self.nan_val = []
        self.pv = None
        self.dtrend = None
        self.device = _lib.default_device() if device is None else int(device)
        kind, gx, gy = _lib.mesh_or_points(GridX, GridY)
        dem = None if GridZ is None else _lib.as_f64(GridZ).reshape(-1)
        self._h = _lib.handle_t()
        _lib.check(_lib.lib().smrfb_idw_create(
            self.npoints, _lib.dptr(self.mx), _lib.dptr(self.my), _lib.dptr(self.mz),
            self.shape[0], self.shape[1], kind, _lib.dptr(gx), _lib.dptr(gy), _lib.dptr(dem),
            float(power), self.device, self._h))

    def kernel_info(self):
        """Which device path serves this handle (smrfb_idw_info)."""
        info = np.zeros(8, dtype=np.int32)
        _lib.check(_lib.lib().smrfb_idw_info(self._h, info.ctypes.data_as(_lib.c_i32_p)))
        return {"far_field": bool(info[0]), "nodes": int(info[1]), "max_near": int(info[2]), "patches": int(info[3]),
                "theta": info[4] / 1000.0, "int8": bool(info[5]), "small": bool(info[6]), "mean_near": info[7] / 1000.0}

    # ------------------------------------------------------------------ batched extension
    def distribute_block(self, data, detrend=False, flag=0, vmin=None, vmax=None, pv=None, out=None, out_dtype=None):
        """Distribute T timesteps at once. data: [T, nsta] (NaN = missing station).

        Returns float64 [T, ny, nx] (page-locked memory from the library's pool unless ``out`` is given); the
        elevation trend coefficients used land in ``self.pv_block``.
        ``pv`` ([T, 2], optional) bypasses the device regression (e.g. np.polyfit coefficients).
        ``out`` may be a preallocated (ideally pinned) C-contiguous [T, ny, nx] array.
        ``out_dtype=np.float32`` stores float32 (the reference's output-file type, contract 1e-5); default float64.
        """
        data = _lib.as_f64(data)
        if data.ndim == 1:
            data = data[None, :]
        T, S = data.shape
        if S != self.npoints:
            raise ValueError("data has %d stations, IDW was built for %d" % (S, self.npoints))
        f32 = self._select_dtype(out_dtype, out)
        out = _lib.out_array(T, self.shape, out, f32)
        pv_in = None if pv is None else _lib.as_f64(pv).reshape(T, 2)
        pv_out = np.empty((T, 2), dtype=np.float64)
        lo = -np.inf if vmin is None else float(vmin)
        hi = np.inf if vmax is None else float(vmax)
        _lib.check(_lib.lib().smrfb_idw_distribute(
            self._h, _lib.dptr(data), T, int(bool(detrend)), int(flag), lo, hi,
            _lib.dptr(pv_in), _lib.dptr(pv_out), _lib.vptr(out)))
        self.pv_block = pv_out
        return out

    # ------------------------------------------------------------------ reference API
    def calculateIDW(self, data, local=False):
        """idw.py:82-94."""
        return self.distribute_block(np.asarray(data, dtype=np.float64)[None, :], detrend=False)[0]

    def detrendedIDW(self, data, flag=0, zeros=None, local=False):
        """idw.py:96-110."""
        data = np.array(data, dtype=np.float64)          # private copy: the caller's row is not touched
        if zeros is None:
            v = self.distribute_block(data[None, :], detrend=True, flag=flag)[0]
            self.pv = self.pv_block[0].copy()
            self.dtrend = data - (self.mz * self.pv[0] + self.pv[1])
            return v
        # rarely used branch (image_data.py always passes zeros=None): the trend is fitted on the
        # original values, the flagged stations then take zeroVal, and negatives are clipped.
        self.detrendData(data, flag, zeros)
        v = self.distribute_block(data[None, :], detrend=True, flag=flag, pv=self.pv[None, :], vmin=0.0)[0]
        return v

    def detrendData(self, data, flag=0, zeros=None):
        """idw.py:112-136: sets ``pv`` and ``dtrend`` (host-side bookkeeping of an [nsta] vector)."""
        ok = ~np.isnan(data)
        pv = np.polyfit(self.mz[ok], data[ok], 1)
        if (flag == 1 and pv[0] < 0) or (flag == -1 and pv[0] > 0):
            pv = np.array([0.0, 0.0])
        self.pv = pv
        if zeros is not None:
            data[zeros] = self.zeroVal
        self.dtrend = data - (self.mz * pv[0] + pv[1])

    def retrendData(self, idw):
        """idw.py:138-144 (the fused kernels do this in their epilogue; kept for API parity)."""
        return idw + self.pv[0] * self.GridZ + self.pv[1]
