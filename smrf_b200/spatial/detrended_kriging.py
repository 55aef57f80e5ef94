"""GPU stand-in for the reference's Cython extension ``smrf.spatial.dk.detrended_kriging``
(detrended_kriging.pyx:32-67): ``call_grid`` with the same arguments and the same in-place output,
so the reference's own ``dk.py`` can run against it unmodified (``from . import detrended_kriging``)."""
import numpy as np

from .. import _lib


def call_grid(ad, dgrid, elevations, weights, nthreads=1):
    """ad [nsta, nsta], dgrid [ngrid, nsta], elevations [nsta], weights [ngrid, nsta] (filled in place)."""
    ad = np.ascontiguousarray(ad, dtype=np.float64)
    dgrid = np.ascontiguousarray(dgrid, dtype=np.float64)
    elevations = np.ascontiguousarray(elevations, dtype=np.float64)
    if not (isinstance(weights, np.ndarray) and weights.dtype == np.float64 and weights.flags["C_CONTIGUOUS"]):
        raise TypeError("weights must be a C-contiguous float64 array (it is written in place)")
    ngrid, nsta = dgrid.shape
    assert ad.shape == (nsta, nsta) and elevations.shape == (nsta,) and weights.shape == (ngrid, nsta)
    _lib.check(_lib.lib().smrfb_krige_grid(nsta, ngrid, _lib.dptr(ad), _lib.dptr(dgrid), _lib.dptr(elevations),
                                           int(nthreads), _lib.dptr(weights)))
    return None
