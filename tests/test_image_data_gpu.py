"""The dispatch layer (mirror of smrf.distribute.image_data) driven the way the reference's variable classes
drive it -- pandas rows, per-variable config, then utils.set_min_max -- against the oracle on RME station data."""
import numpy as np
import pandas as pd
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


class Topo:
    pass


def test_image_data_idw_dk_like_air_temp_and_precip(golden):
    import torch
    assert torch.cuda.is_available()
    from oracle import oracle as orc
    from smrf_b200.distribute import image_data
    from smrf_b200.utils import set_min_max
    g = golden("idw_rme")
    topo = Topo()
    topo.X, topo.Y = np.meshgrid(g["x"], g["y"])
    topo.dem = g["dem"]
    topo.mask = np.ones_like(g["dem"])
    md = pd.DataFrame({"utm_x": g["mx"], "utm_y": g["my"], "elevation": g["mz"]}, index=["RME_176", "RMESP"])
    times = pd.date_range("1998-01-01", periods=len(g["air_temp"]), freq="h")
    ta = pd.DataFrame(g["air_temp"], index=times, columns=md.index)

    var = image_data("air_temp")
    var.getConfig({"distribution": "idw", "detrend": True, "detrend_slope": -1, "idw_power": 2.0, "min": -73.0, "max": 47.0,
                   "stations": None})
    var._initialize(topo, md)
    # per-step, exactly like air_temp.ta.distribute (air_temp.py:73-86)
    var._distribute(ta.iloc[3])
    v = set_min_max(var.air_temp, var.min, var.max)
    assert parity(v, g["out_air_temp"][3], step_scale(g["air_temp"][3])) <= TOL
    # block façade: same numbers served per timestep
    var.prefetch(ta)
    for t in (0, 5, 11):
        var._distribute(ta.iloc[t])
        assert parity(var.air_temp, g["out_air_temp"][t], step_scale(g["air_temp"][t])) <= TOL
    # all-NaN row is rejected like image_data.py:229-231
    bad = ta.iloc[0].copy()
    bad[:] = np.nan
    var._block = None
    with pytest.raises(Exception):
        var._distribute(bad)

    gd = golden("dk_rme")
    topo2 = Topo()
    topo2.X, topo2.Y = np.meshgrid(gd["x"], gd["y"])
    topo2.dem = gd["dem"]
    pr = pd.DataFrame(gd["data"], index=pd.date_range("1998-01-01", periods=len(gd["data"]), freq="h"), columns=md.index)
    pvar = image_data("precip")
    pvar.getConfig({"distribution": "dk", "detrend": True, "detrend_slope": 1, "dk_ncores": 2, "min": 0.0, "max": None, "stations": None})
    pvar._initialize(topo2, md)
    for t in range(3):
        pvar._distribute(pr.iloc[t])
        v = set_min_max(pvar.precip, pvar.min, pvar.max)
        assert parity(v, gd["out"][t], step_scale(gd["data"][t])) <= TOL
