"""Generate tests/golden/grid_local.npz by EXECUTING THE REFERENCE's GRID with grid_local=True.

Run from the repo root:  python tests/golden/make_golden_local.py      (needs /root/reference, read-only)

smrf/spatial/grid.py is read into memory, its two import lines are patched -- scipy.interpolate.interpnd is gone
in SciPy 1.18; smrf.utils.utils needs packages that are absent here, so grid_interpolate_deconstructed
(smrf/utils/utils.py:430-449) is extracted from the reference source in memory -- and exec'd; nothing is written
to disk.  Path pinned: GRID.__init__ neighbour table (grid.py:54-77) and detrendedInterpolationLocal
(grid.py:115-177) with grid_method='linear', flags -1 / 0 / +1.
"""
import os
import re
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SMRF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

from smrf_b200 import synthetic as syn  # noqa: E402


def load_ref_grid_local():
    util_src = open(os.path.join(REF, "smrf/utils/utils.py")).read()
    mt = re.search(r"^def grid_interpolate_deconstructed\(.*?(?=^def )", util_src, re.S | re.M)
    from scipy.interpolate import CloughTocher2DInterpolator, LinearNDInterpolator
    ns = {"np": np, "LinearNDInterpolator": LinearNDInterpolator, "CloughTocher2DInterpolator": CloughTocher2DInterpolator}
    exec(mt.group(0), ns)
    src = open(os.path.join(REF, "smrf/spatial/grid.py")).read()
    src = src.replace("from scipy.interpolate.interpnd import _ndim_coords_from_arrays",
                      "from scipy.interpolate._interpnd import _ndim_coords_from_arrays")
    src = src.replace("from smrf.utils.utils import grid_interpolate_deconstructed", "")
    m = types.ModuleType("ref_grid_local")
    m.__dict__["grid_interpolate_deconstructed"] = ns["grid_interpolate_deconstructed"]
    exec(compile(src, "ref_grid.py", "exec"), m.__dict__)
    return m.GRID


def metadata_for(px, py, pz):
    """HRRR-like metadata frame: utm coordinates, a fake but monotone lat / lon, elevation."""
    idx = ["grid_y%d_x%d" % (i // 100, i % 100) for i in range(len(px))]
    return pd.DataFrame({"utm_x": px, "utm_y": py, "elevation": pz,
                         "latitude": 43.0 + (py - py.min()) / 111000.0,
                         "longitude": -116.0 + (px - px.min()) / 82000.0}, index=idx)


def main():
    GRID = load_ref_grid_local()
    x, y, dem = syn.make_dem(40, 50, dx=100.0, seed=9)
    X, Y = np.meshgrid(x, y)
    px, py, pz = syn.make_source_lattice(x, y, dem, spacing=900.0, jitter=60.0, seed=5)
    keep = ~((px > x[-1] - 1500.0) & (py > y[0] - 1200.0))      # part of the DEM outside the hull
    px, py, pz = px[keep], py[keep], pz[keep]
    fields = syn.make_source_fields(px, py, pz, 4, seed=9)
    md = metadata_for(px, py, pz)
    k = 8
    cfg = {"grid_local": True, "grid_local_n": k, "grid_mask": False}
    ref = GRID(cfg, md.utm_x.values, md.utm_y.values, X, Y, mz=md.elevation.values, GridZ=dem, mask=None, metadata=md)
    nbr = md.index.get_indexer(ref.dist_df.values.ravel()).reshape(len(md), k).astype(np.int32)
    res = {}
    for flag in (-1, 0, 1):
        res["out_f%d" % flag] = np.stack([ref.detrendedInterpolation(pd.Series(r.copy(), index=md.index), flag, "linear")
                                          for r in fields])
    np.savez_compressed(os.path.join(OUT, "grid_local.npz"), px=px, py=py, pz=pz, lat=md.latitude.values,
                        lon=md.longitude.values, x=x, y=y, dem=dem, fields=fields, k=k, nbr=nbr, **res)
    print("grid_local.npz", os.path.getsize(os.path.join(OUT, "grid_local.npz")), "NaN cells per step:",
          int(np.isnan(res["out_f0"][0]).sum()))


if __name__ == "__main__":
    main()
