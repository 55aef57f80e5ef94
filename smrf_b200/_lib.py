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
    "smrfb_grid_distribute_local": (ctypes.c_int, [handle_t, c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                   c_double_p]),
    "smrfb_grid_distribute_local_dev": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                                       ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
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


def default_device():
    return int(os.environ.get("SMRFB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def launch_count():
    return int(lib().smrfb_launch_count())
