"""Generate tests/golden/grid_cubic.npz by EXECUTING THE REFERENCE's GRID with grid_method='cubic'.

Run from the repo root:  python tests/golden/make_golden_cubic.py      (needs /root/reference, read-only)

The reference's smrf/spatial/grid.py is loaded exactly as make_golden_local.py does it (two import lines patched in
memory, grid_interpolate_deconstructed taken from smrf/utils/utils.py:430-449).  Paths pinned, all with 'cubic'
(scipy CloughTocher2DInterpolator; the default grid_method, smrf/framework/CoreConfig.ini:183-186, and what the
gridded gold configs run, smrf/tests/basins/Lakes/gold_hrrr/gold_config.ini:48-55):
  * detrendedInterpolationLocal (grid.py:115-177), flags -1 / 0 / +1        -> local_f{-1,0,1}
  * detrendedInterpolationMask  (grid.py:179-212) with a point mask, flag -1 -> mask_out, mask_pv
  * calculateInterpolation      (grid.py:214-231)                            -> plain_out
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden_local import load_ref_grid_local, metadata_for  # noqa: E402
from smrf_b200 import synthetic as syn  # noqa: E402


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
    res = {}
    for flag in (-1, 0, 1):
        res["local_f%d" % flag] = np.stack([ref.detrendedInterpolation(pd.Series(r.copy(), index=md.index), flag, "cubic")
                                            for r in fields])
    mask = ((X - x.mean()) ** 2 + (Y - y.mean()) ** 2) < 1500.0 ** 2
    cfg2 = {"grid_local": False, "grid_mask": True}
    ref2 = GRID(cfg2, px, py, X, Y, mz=pz, GridZ=dem, mask=mask, metadata=None)
    mo, mpv = [], []
    for r in fields:
        mo.append(ref2.detrendedInterpolation(pd.Series(r.copy()), -1, "cubic"))
        mpv.append(np.array(ref2.pv, dtype=float))
    res["mask_out"] = np.stack(mo)
    res["mask_pv"] = np.stack(mpv)
    res["point_mask"] = np.asarray(ref2.mask, dtype=bool)
    res["plain_out"] = np.stack([ref2.calculateInterpolation(r.copy(), "cubic") for r in fields])
    np.savez_compressed(os.path.join(HERE, "grid_cubic.npz"), px=px, py=py, pz=pz, lat=md.latitude.values,
                        lon=md.longitude.values, x=x, y=y, dem=dem, mask=mask, fields=fields, k=k, **res)
    print("grid_cubic.npz", os.path.getsize(os.path.join(HERE, "grid_cubic.npz")), "NaN cells per step:",
          int(np.isnan(res["plain_out"][0]).sum()), "mask points:", int(res["point_mask"].sum()))


if __name__ == "__main__":
    main()
