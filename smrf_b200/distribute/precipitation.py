"""The distribution calls of smrf.distribute.precipitation.ppt (smrf/distribute/precipitation.py:327-369, 405-425): the
hourly field -- skipped when no station reports precipitation (``data.sum() == 0``) -- and, on the first step of a storm,
the storm total distributed through the SAME interpolator (for DK a second station mask: the cached kriging weights of
both masks stay resident) -- followed, when the precipitation temperature is handed in, by the phase / new-snow density
consumer (Snow.phase_and_density, precipitation.py:376-384, 428-431) on the device.  Storm tracking (envphys.storms) is outside
the path."""
import numpy as np

from ..snow import Snow
from .image_data import image_data as _image_data


class ppt(_image_data):
    variable = 'precip'
    output_variables = {'precip': {'units': 'mm', 'standard_name': 'precipitation_mass', 'long_name': 'Precipitation mass'},
                        'storm_total': {'units': 'mm', 'standard_name': 'precipitation_mass_storm', 'long_name': 'Precipitation mass for the storm period'}}
    post_process_variables = {}

    def __init__(self, pptConfig, start_date=None, time_step=60):
        _image_data.__init__(self, self.variable)
        self.getConfig(pptConfig)
        self.time_step = float(time_step)
        self.start_date = start_date
        self.storms = None                    # DataFrame with 'start', 'end' and one storm-total column per station
        self.storm_id = 0
        self.storming = False
        self.nasde_model = pptConfig.get('new_snow_density_model', 'susong1999') if hasattr(pptConfig, 'get') else 'susong1999'

    def initialize(self, topo, data, date_time=None):
        self.date_time = date_time
        self._initialize(topo, data.metadata)
        self.precip = np.zeros((topo.dem.shape[0], topo.dem.shape[1]))
        self.storm_total = np.zeros((topo.dem.shape[0], topo.dem.shape[1]))
        self.percent_snow = np.zeros(self.precip.shape)
        self.snow_density = np.zeros(self.precip.shape)

    def distribute(self, data, time=None, precip_temp=None):
        # precipitation.py:335-369 (marks2017) / :420-421 (susong1999): the distribution part; with precip_temp (the dew point
        # or wet bulb grid, precipitation.py:271-279) also percent_snow / snow_density (:371-384, :428-431, :394-401)
        if data.sum() > 0.0:
            storm_start = None
            if self.storms is not None:
                for i, s in self.storms.iterrows():
                    if time >= s['start'] and time <= s['end']:
                        self.storm_id, self.storming, storm_start = i, True, s['start']
                        break
                    self.storming = False
            self._distribute(data, zeros=None)                            # clamp [min, max] fused
            if storm_start is not None and time == storm_start:
                storm = self.storms.iloc[self.storm_id]
                self._distribute(storm[self.stations].astype(float), other_attribute='storm_total', clamp=True)
            if precip_temp is not None:
                marks = self.nasde_model == 'marks2017'
                if marks and not (self.storming and np.min(precip_temp) < 2.0):      # precipitation.py:371, 381-384
                    self.snow_density = np.zeros(self.precip.shape)
                    self.percent_snow = np.zeros(self.precip.shape)
                else:
                    self.snow_density, self.percent_snow = Snow.phase_and_density(precip_temp, self.precip, self.nasde_model)
        else:
            self.precip = np.zeros(self.precip.shape)
            if precip_temp is not None:
                self.percent_snow = np.zeros(self.precip.shape)
                self.snow_density = np.zeros(self.precip.shape)

    def distribute_thread(self, smrf_queue, data_queue):
        for date_time in self.date_time:
            data = data_queue[self.variable].get(date_time)
            self.distribute(data, date_time)
            smrf_queue[self.variable].put([date_time, self.precip])
