"""Mirror of the interpolated-stations branch of smrf.distribute.wind.Wind (smrf/distribute/wind/wind.py:161-190): three
_distribute calls per timestep (speed, and the sine / cosine of the station directions), the azimuth of the two
distributed components (arctan2 on the device), min / max on the speed only.  The Winstral and WindNinja wind models
are outside the path (WindNinja's grid_interpolate lives in smrf_b200.utils)."""
import numpy as np

from .image_data import image_data as _image_data
from .. import _lib


class Wind(_image_data):
    variable = 'wind'
    output_variables = {'wind_speed': {'units': 'm/s', 'standard_name': 'wind_speed', 'long_name': 'Wind speed'},
                        'wind_direction': {'units': 'degrees', 'standard_name': 'wind_direction', 'long_name': 'Wind direction'}}
    post_process_variables = {}

    def __init__(self, config):
        _image_data.__init__(self, self.variable)
        self.getConfig(config)

    def initialize(self, topo, data, date_time=None):
        self.date_time = date_time
        self._initialize(topo, data.metadata)

    def distribute(self, data_speed, data_direction):
        # wind.py:163-179
        self._distribute(data_speed, other_attribute='wind_speed', clamp=True)          # wind.py:187-190, fused
        self.u_direction = np.sin(data_direction * np.pi / 180)
        self.v_direction = np.cos(data_direction * np.pi / 180)
        self._distribute(self.u_direction, other_attribute='u_direction_distributed')
        self._distribute(self.v_direction, other_attribute='v_direction_distributed')
        u, v = self.u_direction_distributed, self.v_direction_distributed
        az = _lib.pinned_empty(u.shape, np.float64)
        _lib.check(_lib.lib().smrfb_wind_direction(_lib.dptr(np.ascontiguousarray(u).reshape(-1)), _lib.dptr(np.ascontiguousarray(v).reshape(-1)),
                                                   u.size, _lib.vptr(az), _lib.default_device()))
        self.wind_direction = az

    def distribute_thread(self, smrf_queue, data_queue):
        # wind.py:192-214
        for date_time in self.date_time:
            ws_data = data_queue['wind_speed'].get(date_time)
            wd_data = data_queue['wind_direction'].get(date_time)
            self.distribute(ws_data, wd_data)
            smrf_queue['wind_speed'].put([date_time, self.wind_speed])
            smrf_queue['wind_direction'].put([date_time, self.wind_direction])
