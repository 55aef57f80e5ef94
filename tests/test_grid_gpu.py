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
    with pytest.raises(ValueError):
        grid.calculateInterpolation(g["fields"][0], "quintic")
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


def _local_metadata(g):
    import pandas as pd
    idx = ["p%03d" % i for i in range(len(g["px"]))]
    return pd.DataFrame({"utm_x": g["px"], "utm_y": g["py"], "elevation": g["pz"], "latitude": g["lat"],
                         "longitude": g["lon"]}, index=idx)


def test_grid_local_golden(sp, golden, monkeypatch):
    """grid_local=True, grid_method='linear' (detrendedInterpolationLocal, grid.py:115-177) against the vectors the
    reference produced: neighbour table and values bit-exact (the per-point fits are np.polyfit on the host like the
    reference's, the three interpolations and their combination run in one kernel in griddata's operation order)."""
    import pandas as pd
    g = golden("grid_local")
    X, Y = np.meshgrid(g["x"], g["y"])
    md = _local_metadata(g)
    cfg = {"grid_local": True, "grid_local_n": int(g["k"]), "grid_mask": False}
    grid = sp.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], metadata=md)
    np.testing.assert_array_equal(grid.nbr, g["nbr"])
    for flag in (-1, 0, 1):
        ref = g["out_f%d" % flag]
        for t, row in enumerate(g["fields"]):
            v = grid.detrendedInterpolation(pd.Series(row.copy(), index=md.index), flag, "linear")
            np.testing.assert_array_equal(v, ref[t])                              # bit-exact, NaN outside the hull
        out = grid.distribute_block_local(g["fields"], flag=flag, exact=True)
        np.testing.assert_array_equal(out, ref)
        fast = grid.distribute_block_local(g["fields"], flag=flag)               # closed-form fits, vectorised
        for t in range(len(ref)):
            assert parity(fast[t], ref[t], step_scale(g["fields"][t])) <= TOL
    # clamp fused; one-cell-per-thread kernel (odd cell counts take it) agrees bit for bit
    lo, hi = float(np.nanpercentile(g["out_f0"], 20)), float(np.nanpercentile(g["out_f0"], 80))
    clamped = grid.distribute_block_local(g["fields"], flag=0, vmin=lo, vmax=hi, exact=True)
    np.testing.assert_array_equal(clamped, np.where(np.isnan(g["out_f0"]), np.nan, np.clip(g["out_f0"], lo, hi)))
    monkeypatch.setenv("SMRFB_GRID_ONE_CELL", "1")
    np.testing.assert_array_equal(grid.distribute_block_local(g["fields"], flag=-1, exact=True), g["out_f-1"])
    with pytest.raises(ValueError):
        grid.distribute_block(g["fields"], grid_method="quintic")


def test_grid_local_vs_oracle_larger(sp):
    """A larger synthetic basin (more points than the golden case, 37 steps = two transposition tiles)."""
    import pandas as pd
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(90, 131, dx=60.0, seed=4)
    X, Y = np.meshgrid(x, y)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=700.0, jitter=80.0, seed=8)
    md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                       "longitude": -116.0 + (px - px.min()) / 82000.0})
    f = syn.make_source_fields(px, py, pz, 37, seed=2)
    cfg = {"grid_local": True, "grid_local_n": 12, "grid_mask": False}
    grid = sp.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
    o = orc.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
    out = grid.distribute_block_local(f, flag=-1)
    for t in (0, 17, 36):
        assert parity(out[t], o.detrendedInterpolation(f[t].copy(), -1, "linear"), step_scale(f[t])) <= TOL


def test_grid_local_staged_kernel_matches_direct_loads(sp, monkeypatch):
    """The block path of grid_local runs the folded kernel with the vertex values staged per CTA in shared memory (runs of equal
    simplex along the cell index); with more runs per CTA than fit it loads directly.  Both do the same arithmetic: bit-identical
    to the un-staged kernel (SMRFB_GRID_FOLD_DIRECT) on a coarse source lattice (staged), a fine one (direct fallback inside the
    staged kernel), an odd row length (runs wrap rows), 37 steps (two chunks and a ragged tail), float32 output -- and within
    1e-9 of the oracle."""
    import pandas as pd
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    for (ny, nx, dx, spacing) in ((64, 130, 30.0, 1500.0), (40, 87, 60.0, 400.0), (33, 64, 10.0, 2500.0)):
        x, y, dem = syn.make_dem(ny, nx, dx=dx, seed=6)
        X, Y = np.meshgrid(x, y)
        px, py, pz = syn.make_source_lattice(x, y, dem, spacing=spacing, jitter=spacing / 9.0, seed=3)
        md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                           "longitude": -116.0 + (px - px.min()) / 82000.0})
        f = syn.make_source_fields(px, py, pz, 37, seed=5)
        cfg = {"grid_local": True, "grid_local_n": min(12, len(px)), "grid_mask": False}
        grid = sp.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
        monkeypatch.delenv("SMRFB_GRID_FOLD_DIRECT", raising=False)
        staged = grid.distribute_block_local(f, flag=-1, vmin=-40.0, vmax=40.0)
        staged32 = grid.distribute_block_local(f, flag=-1, vmin=-40.0, vmax=40.0, out_dtype=np.float32)
        monkeypatch.setenv("SMRFB_GRID_FOLD_DIRECT", "1")
        direct = grid.distribute_block_local(f, flag=-1, vmin=-40.0, vmax=40.0)
        monkeypatch.delenv("SMRFB_GRID_FOLD_DIRECT", raising=False)
        np.testing.assert_array_equal(staged, direct)
        np.testing.assert_array_equal(staged32, direct.astype(np.float32))
        o = orc.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
        for t in (0, 15, 16, 36):
            ref = orc.set_min_max(o.detrendedInterpolation(f[t].copy(), -1, "linear"), -40.0, 40.0)
            assert parity(staged[t], ref, step_scale(f[t])) <= TOL
        # the Clough-Tocher kernel stages the same way (8 steps at a time): local detrend (two fields) and mask detrend (one)
        f11 = f[:11]
        cub = {}
        for tag in ("staged", "direct"):
            if tag == "direct":
                monkeypatch.setenv("SMRFB_GRID_CUBIC_DIRECT", "1")
            cub[tag] = (grid.distribute_block_local(f11, flag=-1, grid_method="cubic", vmin=-40.0, vmax=40.0),
                        grid.distribute_block(f11, detrend=True, flag=-1, grid_method="cubic"))
            monkeypatch.delenv("SMRFB_GRID_CUBIC_DIRECT", raising=False)
        np.testing.assert_array_equal(cub["staged"][0], cub["direct"][0])
        np.testing.assert_array_equal(cub["staged"][1], cub["direct"][1])
        ref = orc.set_min_max(o.detrendedInterpolation(f11[10].copy(), -1, "cubic"), -40.0, 40.0)
        assert parity(cub["staged"][0][10], ref, step_scale(f11[10])) <= TOL


def _md(g):
    import pandas as pd
    return pd.DataFrame({"utm_x": g["px"], "utm_y": g["py"], "elevation": g["pz"], "latitude": g["lat"], "longitude": g["lon"]})


def test_grid_cubic_golden(sp, golden):
    """grid_method='cubic' (scipy CloughTocher2DInterpolator; the reference's default) against vectors produced by
    executing the reference (tests/golden/make_golden_cubic.py): local detrend (grid.py:115-177 + utils.py:447), mask
    detrend (grid.py:179-212) and plain interpolation (grid.py:214-231).  1e-9 relative, identical NaN pattern."""
    import pandas as pd
    g = golden("grid_cubic")
    X, Y = np.meshgrid(g["x"], g["y"])
    md = _md(g)
    cfg = {"grid_local": True, "grid_local_n": int(g["k"]), "grid_mask": False}
    grid = sp.GRID(cfg, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], metadata=md)
    for flag in (-1, 0, 1):
        ref = g["local_f%d" % flag]
        for t, row in enumerate(g["fields"]):
            v = grid.detrendedInterpolation(pd.Series(row.copy(), index=md.index), flag, "cubic")
            assert parity(v, ref[t], step_scale(row)) <= TOL, (flag, t)
        blk = grid.distribute_block_local(g["fields"], flag=flag, grid_method="cubic")          # closed-form fits, one launch
        for t in range(len(ref)):
            assert parity(blk[t], ref[t], step_scale(g["fields"][t])) <= TOL, (flag, t)
    grid2 = sp.GRID({"grid_local": False, "grid_mask": True}, g["px"], g["py"], X, Y, mz=g["pz"], GridZ=g["dem"], mask=g["mask"])
    np.testing.assert_array_equal(grid2.mask, g["point_mask"])
    for t, row in enumerate(g["fields"]):
        v = grid2.detrendedInterpolation(pd.Series(row.copy()), -1, "cubic")
        np.testing.assert_array_equal(grid2.pv, g["mask_pv"][t])                   # np.polyfit on the host, as the reference
        assert parity(v, g["mask_out"][t], step_scale(row)) <= TOL, t
        assert parity(grid2.calculateInterpolation(row.copy(), "cubic"), g["plain_out"][t], step_scale(row)) <= TOL, t
    blk = grid2.distribute_block(g["fields"], detrend=True, flag=-1, grid_method="cubic")
    for t in range(len(g["fields"])):
        assert parity(blk[t], g["mask_out"][t], step_scale(g["fields"][t])) <= TOL
    # clamp fused, float32 output mode (contract 1e-5)
    lo, hi = float(np.nanpercentile(g["plain_out"], 20)), float(np.nanpercentile(g["plain_out"], 80))
    cl = grid2.distribute_block(g["fields"], grid_method="cubic", vmin=lo, vmax=hi)
    want = np.where(np.isnan(g["plain_out"]), np.nan, np.clip(g["plain_out"], lo, hi))
    for t in range(len(want)):
        assert parity(cl[t], want[t], step_scale(g["fields"][t])) <= TOL
    f32 = grid2.distribute_block(g["fields"], grid_method="cubic", out_dtype=np.float32)
    assert f32.dtype == np.float32
    for t in range(len(want)):
        assert parity(f32[t], g["plain_out"][t], step_scale(g["fields"][t])) <= 1e-5


def test_grid_cubic_vs_oracle_larger(sp):
    """Larger synthetic basin, 37 steps (several gradient-estimation threads), against the oracle (SciPy called where
    the reference calls it)."""
    import pandas as pd
    from oracle import oracle as orc
    from smrf_b200 import synthetic as syn
    x, y, dem = syn.make_dem(90, 131, dx=60.0, seed=4)
    X, Y = np.meshgrid(x, y)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=700.0, jitter=80.0, seed=8)
    md = pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz, "latitude": 43.0 + (py - py.min()) / 111000.0,
                       "longitude": -116.0 + (px - px.min()) / 82000.0})
    f = syn.make_source_fields(px, py, pz, 37, seed=2)
    cfg = {"grid_local": True, "grid_local_n": 12, "grid_mask": False}
    grid = sp.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
    o = orc.GRID(cfg, px, py, X, Y, mz=pz, GridZ=dem, metadata=md)
    out = grid.distribute_block_local(f, flag=-1, grid_method="cubic")
    for t in (0, 17, 36):
        assert parity(out[t], o.detrendedInterpolation(f[t].copy(), -1, "cubic"), step_scale(f[t])) <= TOL
    cfg2 = {"grid_local": False, "grid_mask": False}
    g2 = sp.GRID(cfg2, px, py, X, Y, mz=pz, GridZ=dem)
    o2 = orc.GRID(cfg2, px, py, X, Y, mz=pz, GridZ=dem)
    out = g2.distribute_block(f, detrend=True, flag=-1, grid_method="cubic", vmin=-5.0, vmax=5.0)
    for t in (0, 20, 36):
        ref = orc.set_min_max(o2.detrendedInterpolation(f[t].copy(), -1, "cubic"), -5.0, 5.0)
        assert parity(out[t], ref, step_scale(f[t])) <= TOL
