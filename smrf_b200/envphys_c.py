"""Drop-in for the two entry points of ``smrf.envphys.core.envphys_c`` that consume a freshly distributed grid
(SURVEY.md 8f N-a): ``cdewpt`` (dewpt.c:14-229) and ``cwbt`` (iwbt.c:84-134).  Both run on the device (smrfb_dew_point,
smrfb_wet_bulb); no CPU fallback.
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


def cwbt(ta, td, z, tw, tolerance=0, nthreads=1, device=None):
    """envphys_c.pyx cwbt(ta, td, z, tw, tolerance, nthreads): wet / ice bulb temperature (degC) from air temperature, dew point
    (degC) and elevation (m), written into ``tw`` in place.  Where the reference prints and calls exit(-1) (a temperature below
    0 K, no convergence in 10 Newton steps) this raises SmrfbError; ``tolerance`` must be positive (with the reference's default 0
    the iteration can only end through that exit)."""
    ta, td, z = _lib.as_f64(ta), _lib.as_f64(td), _lib.as_f64(z)
    if not (ta.size == td.size == z.size):
        raise ValueError("ta, td and z must have the same size")
    if not (isinstance(tw, np.ndarray) and tw.dtype == np.float64 and tw.flags["C_CONTIGUOUS"] and tw.size == ta.size):
        raise TypeError("tw must be a C-contiguous float64 array of ta's size (it is filled in place)")
    dev = _lib.default_device() if device is None else device
    _lib.check(_lib.lib().smrfb_wet_bulb(_lib.dptr(ta.reshape(-1)), _lib.dptr(td.reshape(-1)), _lib.dptr(z.reshape(-1)), ta.size,
                                         float(tolerance), _lib.dptr(tw.reshape(-1)), dev))
    return None
