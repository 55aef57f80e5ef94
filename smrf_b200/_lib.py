"""ctypes binding of libsmrfb.so (include/smrfb.h).  There is no CPU fallback: if the CUDA
library is missing or a call fails, this raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmrfb.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
handle_t = ctypes.c_void_p

_lib = None


class SmrfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsmrfb error %d: %s" % (code, msg))
        self.code = code


_SIGS = {
    "smrfb_last_error": (ctypes.c_char_p, []),
    "smrfb_version": (ctypes.c_int, []),
    "smrfb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "smrfb_destroy": (ctypes.c_int, [handle_t]),
    "smrfb_synchronize": (ctypes.c_int, [handle_t]),
    "smrfb_launch_count": (ctypes.c_int64, []),
    "smrfb_set_output_dtype": (ctypes.c_int, [handle_t, ctypes.c_int]),
    "smrfb_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "smrfb_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "smrfb_idw_create": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, ctypes.c_int,
                                        ctypes.POINTER(handle_t)]),
    "smrfb_idw_distribute": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "smrfb_idw_distribute_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_idw_distribute_window_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                       ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_idw_info": (ctypes.c_int, [handle_t, c_i32_p]),
    "smrfb_idw_far_timing": (ctypes.c_int, [handle_t, ctypes.c_int, c_double_p]),
    "smrfb_dk_distribute_window_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, c_double_p, ctypes.c_int, ctypes.c_double,
                                                      ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_dk_create": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(handle_t)]),
    "smrfb_dk_distribute": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                           c_double_p, c_double_p, c_double_p]),
    "smrfb_dk_distribute_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, c_double_p, ctypes.c_int, ctypes.c_double,
                                               ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p]),
    "smrfb_dk_weights": (ctypes.c_int, [handle_t, c_u8_p, c_double_p]),
    "smrfb_dk_stats": (ctypes.c_int, [handle_t, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_int)]),
    "smrfb_krige_grid": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int,
                                        c_double_p]),
    "smrfb_grid_create": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, c_i32_p,
                                         c_double_p, c_i32_p, c_u8_p, ctypes.c_int, ctypes.POINTER(handle_t)]),
    "smrfb_grid_set_nearest": (ctypes.c_int, [handle_t, c_i32_p]),
    "smrfb_grid_distribute": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_double, c_double_p, c_double_p, c_double_p]),
    "smrfb_grid_distribute_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_grid_set_local_exact": (ctypes.c_int, [handle_t, ctypes.c_int]),
    "smrfb_grid_distribute_local": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                   c_double_p]),
    "smrfb_grid_distribute_local_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                                       ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_grid_set_cubic": (ctypes.c_int, [handle_t, c_double_p]),
    "smrfb_grid_distribute_cubic": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_int, c_double_p,
                                                   ctypes.c_double, ctypes.c_double, c_double_p]),
    "smrfb_grid_distribute_cubic_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                       ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_dew_point": (ctypes.c_int, [c_double_p, ctypes.c_int64, ctypes.c_double, c_double_p, ctypes.c_int]),
    "smrfb_dew_point_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_snow_phase": (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int64, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int]),
    "smrfb_snow_phase_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_wet_bulb": (ctypes.c_int, [c_double_p, c_double_p, c_double_p, ctypes.c_int64, ctypes.c_double, c_double_p, ctypes.c_int]),
    "smrfb_wet_bulb_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_set_consumer": (ctypes.c_int, [handle_t, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]),
    "smrfb_wind_direction": (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int]),
    "smrfb_wind_direction_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_interp_create": (ctypes.c_int, [ctypes.c_int64, c_i32_p, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(handle_t)]),
    "smrfb_interp_apply": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_double, c_double_p]),
    "smrfb_interp_apply_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "smrfb_set_min_max": (ctypes.c_int, [c_double_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
}


def lib():
    """Load libsmrfb.so (once). Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m smrf_b200.build` "
                               "(this package has no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SmrfbError(rc, lib().smrfb_last_error().decode("utf-8", "replace"))


def dptr(a):
    """double* of a C-contiguous float64 array (or NULL for None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need C-contiguous float64"
    return a.ctypes.data_as(c_double_p)


def as_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def mesh_or_points(GridX, GridY):
    """Detect np.meshgrid(x, y) structure (smrf/data/load_topo.py:65). Returns (kind, gx, gy)."""
    GridX = np.asarray(GridX, dtype=np.float64)
    GridY = np.asarray(GridY, dtype=np.float64)
    assert GridX.shape == GridY.shape and GridX.ndim == 2
    x = GridX[0, :]
    y = GridY[:, 0]
    if np.array_equal(GridX, np.broadcast_to(x[None, :], GridX.shape)) and \
            np.array_equal(GridY, np.broadcast_to(y[:, None], GridY.shape)):
        return 0, np.ascontiguousarray(x), np.ascontiguousarray(y)
    return 1, np.ascontiguousarray(GridX.ravel()), np.ascontiguousarray(GridY.ravel())


class _PinnedPool:
    """Page-locked output blocks (smrfb_host_alloc), recycled by size: a per-timestep run loop gets its [ny, nx]
    result in pinned memory (full-rate D2H) without paying cudaHostAlloc on every call -- the array of step n-1 is
    usually collected before step n+1 asks for a block."""

    def __init__(self, max_free_bytes=8 << 30):
        import threading
        self.free = {}
        self.free_bytes = 0
        self.max_free_bytes = max_free_bytes
        self.lock = threading.Lock()

    def take(self, nbytes):
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                self.free_bytes -= nbytes
                return lst.pop()
        p = ctypes.c_void_p()
        check(lib().smrfb_host_alloc(nbytes, ctypes.byref(p)))
        return p.value

    def give(self, ptr, nbytes):
        with self.lock:
            if self.free_bytes + nbytes <= self.max_free_bytes:
                self.free.setdefault(nbytes, []).append(ptr)
                self.free_bytes += nbytes
                return
        try:
            lib().smrfb_host_free(ptr)
        except Exception:
            pass


_pool = _PinnedPool()


def pinned_empty(shape, dtype=np.float64):
    """np.empty(shape, dtype) in page-locked memory from the pool; the block returns to the pool when the array
    (and every view of it) has been collected.  The array is an ordinary C-contiguous ndarray owned by Python."""
    import weakref
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(n * dtype.itemsize, 1)
    ptr = _pool.take(nbytes)
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(buf, _pool.give, ptr, nbytes)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def out_array(T, shape, out, f32):
    """The [T, ny, nx] result array of a distribute call: the caller's, or a pinned one."""
    dt = np.float32 if f32 else np.float64
    if out is None:
        return pinned_empty((T,) + tuple(shape), dt)
    if not (isinstance(out, np.ndarray) and out.dtype == dt and out.flags["C_CONTIGUOUS"] and out.shape == (T,) + tuple(shape)):
        raise ValueError("out must be a C-contiguous %s array of shape %r" % (np.dtype(dt).name, (T,) + tuple(shape)))
    return out


def vptr(a):
    return a.ctypes.data_as(c_double_p)


class HandleMixin:
    """Shared by the drop-in classes: handle lifetime and the output element type of the next call."""
    _h = None
    _f32 = False

    def _select_dtype(self, out_dtype, out):
        dt = out_dtype if out_dtype is not None else (out.dtype if out is not None else np.float64)
        dt = np.dtype(dt)
        if dt not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("out_dtype must be float64 (parity contract 1e-9) or float32 (1e-5), got %s" % dt)
        f32 = dt == np.dtype(np.float32)
        if f32 != self._f32:
            check(lib().smrfb_set_output_dtype(self._h, int(f32)))
            self._f32 = f32
        return f32

    def dew_point_of(self, call, T, tol=0.01, out_dtype=np.float64):
        """Run ``call()`` -- one distribute_block / distribute_block_local call on THIS handle, T steps, element type
        ``out_dtype`` -- with the dew-point consumer hooked on (smrfb_set_consumer): returns (distributed block, dew point
        block), the second computed on the device from the first while it is still in the handle's output ring
        (vapor_pressure.py:136-141 without re-reading the grid)."""
        dt = np.dtype(out_dtype)
        out2 = pinned_empty((T,) + tuple(self.shape), dt)
        check(lib().smrfb_set_consumer(self._h, 1, float(tol), out2.ctypes.data_as(ctypes.c_void_p)))
        try:
            out = call()
        finally:
            check(lib().smrfb_set_consumer(self._h, 0, 0.0, None))
        if out.dtype != dt:
            raise ValueError("dew_point_of: the call produced %s, out_dtype says %s" % (out.dtype, dt))
        return out, out2

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().smrfb_destroy(h)
            except Exception:
                pass
            self._h = None


def default_device():
    return int(os.environ.get("SMRFB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def launch_count():
    return int(lib().smrfb_launch_count())
