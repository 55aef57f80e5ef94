"""Mirror of smrf.distribute.cloud_factor.cf (smrf/distribute/cloud_factor.py:40-104) on the CUDA-backed image_data."""
from .image_data import image_data as _image_data


class cf(_image_data):
    variable = 'cloud_factor'
    output_variables = {'cloud_factor': {'units': 'None', 'standard_name': 'cloud_factor', 'long_name': 'cloud factor'}}
    post_process_variables = {}

    def __init__(self, config):
        _image_data.__init__(self, self.variable)
        self.getConfig(config)

    def initialize(self, topo, data, date_time=None):
        self.date_time = date_time
        self._initialize(topo, data.metadata)

    def distribute(self, data):
        # cloud_factor.py:72-86 (clamp to [min, max] fused into the store)
        self._distribute(data)

    def distribute_thread(self, smrf_queue, data_queue):
        for date_time in self.date_time:
            data = data_queue[self.variable].get(date_time)
            self.distribute(data)
            smrf_queue[self.variable].put([date_time, self.cloud_factor])
