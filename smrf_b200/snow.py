"""Drop-in for ``smrf.envphys.snow.Snow`` (smrf/envphys/snow.py:28-75): precipitation phase and new-snow density from the
precipitation temperature and the distributed precipitation, on the device (smrfb_snow_phase; no CPU fallback)."""
import numpy as np

from . import _lib


class Snow:
    MODELS = {"susong1999": 0, "piecewise_susong1999": 1, "marks2017": 2}      # nasde_model/{susong_1999,piecewise_suosong_1999,marks_2017}.py

    @staticmethod
    def phase_and_density(temperature, precipitation, nasde_model, device=None):
        """snow.py:37-75: returns (snow_density [kg/m^3], percent_snow [0..1]), new arrays of ``temperature``'s shape; the
        inputs are not modified (the reference copies them before its models clip the temperature in place)."""
        if nasde_model not in Snow.MODELS:
            raise ValueError("{0} is not an implemented new accumulated snow density (NASDE) model! Check the config file "
                             "under precipitation".format(nasde_model))
        t = _lib.as_f64(np.asarray(temperature, dtype=np.float64))
        pp = _lib.as_f64(np.asarray(precipitation, dtype=np.float64))
        if t.shape != pp.shape:
            raise ValueError("temperature and precipitation must have the same shape")
        pcs = np.empty(t.shape)
        rho = np.empty(t.shape)
        dev = _lib.default_device() if device is None else device
        _lib.check(_lib.lib().smrfb_snow_phase(_lib.dptr(t.reshape(-1)), _lib.dptr(pp.reshape(-1)), t.size, Snow.MODELS[nasde_model],
                                               _lib.dptr(pcs.reshape(-1)), _lib.dptr(rho.reshape(-1)), dev))
        return rho, pcs
