"""GPU parity of GRID-linear (through the C-ABI) against the reference-generated golden vectors
(griddata executed by the reference's grid.py) -- bit-exact values, simplex selection and point
mask -- and against the oracle at a larger size."""
import numpy as np
import pytest

from util import TOL, parity, step_scale

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import smrf_b200.spatial as sp
    return sp


def test_grid_golden(sp, golden):
    g = golden("grid_synth")
    X, Y = np.meshgrid(g["x"], g["y"])
    for gm in (1, 0):
        cfg = {"grid_local": False, "grid_local_n": 25, "grid_mask": bool(gm)}
        grid = sp.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], mask=g["mask"])
        np.testing.assert_array_equal(grid.mask, g["ptmask_%d" % gm])                  # bit-exact point mask
        for flag in (-1, 0):
            out = grid.distribute_block(g["fields"], detrend=True, flag=flag)
            ref = g["out_m%d_f%d" % (gm, flag)]
            np.testing.assert_allclose(grid.pv_block, g["pv_m%d_f%d" % (gm, flag)], rtol=1e-9, atol=1e-12)
            for t in range(len(ref)):
                assert parity(out[t], ref[t], step_scale(g["fields"][t])) <= TOL
            # with the reference's own polyfit coefficients fed in, the result is bit-identical to griddata's
            out2 = grid.distribute_block(g["fields"], detrend=True, flag=flag, pv=g["pv_m%d_f%d" % (gm, flag)])
            np.testing.assert_array_equal(out2, ref)
    out = grid.distribute_block(g["fields"], detrend=False)
    np.testing.assert_array_equal(out, g["out_plain"])                                  # bit-exact, incl. NaN outside hull
    v = grid.calculateInterpolation(g["fields"][2], "linear")
    np.testing.assert_array_equal(v, g["out_plain"][2])
    import pandas as pd
    v = grid.detrendedInterpolation(pd.Series(g["fields"][1]), -1, "linear")
    np.testing.assert_array_equal(np.isnan(v), np.isnan(g["out_m0_f-1"][1]))
    with pytest.raises(NotImplementedError):
        grid.calculateInterpolation(g["fields"][0], "cubic")
    # grid_method 'nearest' (NearestNDInterpolator): bit-exact copy of the nearest source value, detrended form within 1e-9
    out = grid.distribute_block(g["fields"], detrend=False, grid_method="nearest")
    np.testing.assert_array_equal(out, g["out_plain_nearest"])
    for t in range(len(g["fields"])):
        v = grid.detrendedInterpolation(g["fields"][t], -1, "nearest")
        assert parity(v, g["out_m0_f-1_nearest"][t], step_scale(g["fields"][t])) <= TOL


def test_grid_vs_oracle_larger(sp):
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(301, 403, dx=10.0, seed=31)
    X, Y = np.meshgrid(x, y)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=700.0, jitter=80.0, seed=5)
    fields = syn.make_source_fields(px, py, pz, 37, seed=9)
    mask = ((X - x.mean()) ** 2 + (Y - y.mean()) ** 2) < 1200.0 ** 2
    cfg = {"grid_local": False, "grid_mask": True}
    grid = sp.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, mask=mask)
    o = orc.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, mask=mask)
    np.testing.assert_array_equal(grid.mask, o.mask)
    plan = orc.linear_plan(px, py, X, Y)
    np.testing.assert_array_equal(grid.cell_simplex, plan["simplex"])                 # bit-exact simplex selection
    out = grid.distribute_block(fields, detrend=True, flag=-1, vmin=-20.0, vmax=8.0)
    for t in (0, 5, 36):
        ref = orc.set_min_max(o.detrendedInterpolation(fields[t].copy(), -1, "linear"), -20.0, 8.0)
        assert parity(out[t], ref, step_scale(fields[t])) <= TOL
        # decomposed oracle with the device's coefficients: identical bits
        pv = grid.pv_block[t]
        ref2 = orc.linear_apply(plan, fields[t] - (pz * pv[0] + pv[1]), X, Y) + pv[0] * dem + pv[1]
        np.testing.assert_array_equal(out[t], orc.set_min_max(ref2, -20.0, 8.0))
    # NaN at a source point propagates to exactly the cells of the simplices that use it
    f = fields[:2].copy()
    f[1, 40] = np.nan
    out = grid.distribute_block(f, detrend=False)
    uses = np.any(plan["simplices"][np.maximum(plan["simplex"], 0)] == 40, axis=1) & (plan["simplex"] >= 0)
    np.testing.assert_array_equal(np.isnan(out[1]).ravel(), uses | (plan["simplex"] < 0))
