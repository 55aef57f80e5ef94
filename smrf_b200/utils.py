"""Host-facing helpers with the reference's names (smrf/utils/utils.py)."""
import numpy as np

from . import _lib


def set_min_max(data, min_val, max_val, device=None):
    """NaN-preserving clamp, in place, same contract as smrf/utils/utils.py:137-161.

    The distribute calls of this package already fuse the clamp (pass min/max there); this
    stand-alone form exists so `utils.set_min_max(grid, min, max)` call sites keep working.
    It runs on the GPU (smrfb_set_min_max); there is no NumPy fallback.
    """
    if max_val is None:
        max_val = np.inf
    if min_val is None:
        min_val = -np.inf
    if not (isinstance(data, np.ndarray) and data.dtype == np.float64 and data.flags["C_CONTIGUOUS"]):
        raise TypeError("set_min_max needs a C-contiguous float64 ndarray (it is modified in place)")
    dev = _lib.default_device() if device is None else device
    _lib.check(_lib.lib().smrfb_set_min_max(_lib.dptr(data), data.size, float(min_val), float(max_val), dev))
    return data
