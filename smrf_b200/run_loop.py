"""Run-loop façade (SURVEY.md 8f N-d): drive the variable classes the way smrf.framework.model_framework does --
``initialize`` hands every variable the whole data object (model_framework.py:362-363), the threaded loop starts one
Python thread per variable running its ``distribute_thread`` over date-indexed queues (model_framework.py:568-632,
smrf/utils/queue.py) -- and put the block façade underneath: ``prefetch_blocks`` distributes the next block of timesteps of
every variable in one fused launch each, so the per-timestep ``distribute`` calls of an UNCHANGED loop are served from
page-locked blocks instead of launching once per step."""
import threading

import pandas as pd


class DateQueue:
    """Minimal date-indexed blocking queue with the put([date_time, data]) / get(date_time) protocol of
    smrf.utils.queue.DateQueue_Threading (queue.py:12-120)."""

    def __init__(self, timeout=600.0):
        self._d = {}
        self._cv = threading.Condition()
        self.timeout = timeout

    def put(self, item):
        date_time, data = item
        with self._cv:
            self._d[date_time] = data
            self._cv.notify_all()

    def get(self, date_time, remove=False):
        with self._cv:
            if not self._cv.wait_for(lambda: date_time in self._d, timeout=self.timeout):
                raise TimeoutError("DateQueue: nothing arrived for %s" % (date_time,))
            return self._d.pop(date_time) if remove else self._d[date_time]

    def clean(self, date_time):
        with self._cv:
            self._d.pop(date_time, None)


def data_queues(frames):
    """{variable: DataFrame [time x station]} -> {variable: DateQueue of rows}, as model_framework.py:575-590 fills them."""
    qs = {}
    for name, frame in frames.items():
        q = DateQueue()
        for t, row in frame.iterrows():
            q.put([t, row])
        qs[name] = q
    return qs


def prefetch_blocks(variables, frames, t0, t1):
    """Distribute timesteps [t0, t1) of every variable now (one fused launch per variable and field).  ``variables`` maps a
    name to an image_data-derived object; wind gets its three fields (speed clamped, direction components not)."""
    import numpy as np
    for name, var in variables.items():
        if name == "wind":
            sp, di = frames["wind_speed"].iloc[t0:t1], frames["wind_direction"].iloc[t0:t1]
            var.prefetch(sp, other_attribute="wind_speed", clamp=True)
            var.prefetch(np.sin(di * np.pi / 180), other_attribute="u_direction_distributed", clamp=False)
            var.prefetch(np.cos(di * np.pi / 180), other_attribute="v_direction_distributed", clamp=False)
        else:
            var.prefetch(frames[name].iloc[t0:t1])


def run_threaded(variables, frames, date_time, outputs=None, block_steps=None, sink=None):
    """The threaded distribute loop of model_framework.py:568-632 for the variables on this path.

    variables: {'air_temp': ta, 'vapor_pressure': vp, 'wind': Wind, 'precip': ppt, 'cloud_factor': cf} (any subset;
    vapor_pressure needs air_temp).  frames: {name: DataFrame}.  block_steps: prefetch that many timesteps at a time
    (None: per-step launches, like the reference).  sink(name, date_time, grid) is called from an output thread for every
    name in ``outputs``.  Returns {name: DateQueue} of the distributed grids."""
    date_time = list(date_time)
    out_names = {"air_temp": ["air_temp"], "vapor_pressure": ["vapor_pressure", "dew_point", "precip_temp"],
                 "wind": ["wind_speed", "wind_direction"], "precip": ["precip"], "cloud_factor": ["cloud_factor"]}
    smrf_queue = {n: DateQueue() for v in variables for n in out_names[v]}
    dq = data_queues(frames)
    for var in variables.values():
        var.date_time = date_time
    nb = block_steps or len(date_time)
    errors = []

    def guard(fn, *a):
        try:
            fn(*a)
        except Exception as e:       # noqa: BLE001 -- reported to the caller below
            errors.append(e)

    threads = []
    for b0 in range(0, len(date_time), nb):
        if block_steps:
            prefetch_blocks(variables, frames, b0, min(b0 + nb, len(date_time)))
        for var in variables.values():
            var.date_time = date_time[b0:b0 + nb]
        threads = [threading.Thread(target=guard, args=(var.distribute_thread, smrf_queue, dq), name=name)
                   for name, var in variables.items()]
        if sink is not None:
            def drain(ts=tuple(date_time[b0:b0 + nb])):
                for t in ts:
                    for n in (outputs or list(smrf_queue)):
                        sink(n, t, smrf_queue[n].get(t))
            threads.append(threading.Thread(target=guard, args=(drain,), name="output"))
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    return smrf_queue
