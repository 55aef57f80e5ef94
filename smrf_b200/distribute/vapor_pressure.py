"""Mirror of smrf.distribute.vapor_pressure.vp (smrf/distribute/vapor_pressure.py:60-190) on the CUDA-backed image_data:
distribute vapor pressure, clamp, dew point (dewpt.c through the device consumer: the freshly distributed grid is not
re-read from host memory), ``dpt >= ta -> ta - 0.2``.  The wet-bulb branch (iwbt.c) is outside the path."""
import numpy as np

from .image_data import image_data as _image_data


class vp(_image_data):
    variable = 'vapor_pressure'
    output_variables = {'vapor_pressure': {'units': 'pascal', 'standard_name': 'vapor_pressure', 'long_name': 'Vapor pressure'},
                        'dew_point': {'units': 'degree_Celsius', 'standard_name': 'dew_point_temperature', 'long_name': 'Dew point temperature'},
                        'precip_temp': {'units': 'degree_Celsius', 'standard_name': 'precip_temperature', 'long_name': 'Precip temperature'}}
    post_process_variables = {}

    def __init__(self, vpConfig, precip_temp_method='dew_point'):
        _image_data.__init__(self, self.variable)
        self.getConfig(vpConfig)
        self.precip_temp_method = precip_temp_method
        if precip_temp_method != 'dew_point':
            raise NotImplementedError("precip_temp_method %r: only 'dew_point' is on this path" % (precip_temp_method,))

    def initialize(self, topo, data, date_time=None):
        # vapor_pressure.py:80-102
        self.date_time = date_time
        self._initialize(topo, data.metadata)
        self.dem = topo.dem

    def _interp(self):
        return getattr(self, self.config["distribution"])

    def prefetch(self, frame, other_attribute=None, clamp=True):
        """Block façade: vapor pressure AND its dew point for every row of ``frame`` in one pass over device memory."""
        if other_attribute is not None:
            return _image_data.prefetch(self, frame, other_attribute, clamp)
        tol = self.config.get('dew_point_tolerance', 0.01)
        blk, dpt = self._interp().dew_point_of(lambda: _image_data.prefetch(self, frame), len(frame), tol)
        self._dpt_block = dpt
        return blk

    def distribute(self, data, ta):
        # vapor_pressure.py:104-164
        tol = self.config.get('dew_point_tolerance', 0.01)
        h = self._interp()
        got = {}

        def call():
            self._distribute(data)
            got["vp"] = self.vapor_pressure
            return self.vapor_pressure[None] if self.vapor_pressure.ndim == 2 else self.vapor_pressure
        key = getattr(data, "name", None)
        if self._block is not None and key in self._block_index:      # served from a prefetched block (dew point included)
            self._distribute(data)
            dpt = self._dpt_block[self._block_index[key]]
        else:
            _, dpt = h.dew_point_of(call, 1, tol)
            dpt = dpt[0]
        ind = dpt >= ta                                               # vapor_pressure.py:143-147
        if np.sum(ind) > 0:
            dpt[ind] = ta[ind] - 0.2
        self.dew_point = dpt
        self.precip_temp = dpt

    def distribute_thread(self, smrf_queue, data_queue):
        # vapor_pressure.py:166-190
        for date_time in self.date_time:
            data = data_queue[self.variable].get(date_time)
            ta = smrf_queue['air_temp'].get(date_time)
            self.distribute(data, ta)
            smrf_queue[self.variable].put([date_time, self.vapor_pressure])
            smrf_queue['dew_point'].put([date_time, self.dew_point])
            smrf_queue['precip_temp'].put([date_time, self.precip_temp])
