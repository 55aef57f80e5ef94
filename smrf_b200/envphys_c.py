"""Drop-in for the two entry points of ``smrf.envphys.core.envphys_c`` that consume a freshly distributed grid
(SURVEY.md 8f N-a): ``cdewpt`` (dewpt.c:14-229).  Runs on the device (smrfb_dew_point); no CPU fallback.
The fused form -- dew point computed while the vapor-pressure block is still in device memory -- is
``distribute_block(..., dew_point=True)`` of the interpolators, used by ``smrf_b200.distribute.vapor_pressure.vp``."""
import numpy as np

from . import _lib


def cdewpt(vp, dwpt, tolerance=0, nthreads=1, device=None):
    """envphys_c.pyx cdewpt(vp, dwpt, tolerance, nthreads): dew point (degC) of ``vp`` (Pa) written into ``dwpt`` in place."""
    vp = _lib.as_f64(vp)
    if not (isinstance(dwpt, np.ndarray) and dwpt.dtype == np.float64 and dwpt.flags["C_CONTIGUOUS"] and dwpt.size == vp.size):
        raise TypeError("dwpt must be a C-contiguous float64 array of vp's size (it is filled in place)")
    dev = _lib.default_device() if device is None else device
    _lib.check(_lib.lib().smrfb_dew_point(_lib.dptr(vp.reshape(-1)), vp.size, float(tolerance), _lib.dptr(dwpt.reshape(-1)), dev))
    return None
