"""Quick device-resident timing of the IDW kernel (development aid, not the bench)."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from smrf_b200 import _lib, synthetic as syn  # noqa: E402


def main():
    ny = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 128
    x, y, dem = syn.make_dem(ny, nx, dx=30.0, noise=False)
    mx, my, mz = syn.make_stations(S, x, y, dem)
    import os
    if os.environ.get("DROP_MODE") == "block":       # the bench's outage model: a dozen distinct NaN masks per batch
        data, nm = syn.dropouts_block(syn.make_series("air_temp", mz, T), max_masks=12, per_year=40.0)
        print("block outages:", nm, "distinct masks, %.2f %% missing" % (100 * np.isnan(data).mean()))
    else:
        data = syn.dropouts_iid(syn.make_series("air_temp", mz, T), p=float(os.environ.get("DROP_P", "0.02")))
    L = _lib.lib()
    h = _lib.handle_t()
    _lib.check(L.smrfb_idw_create(S, _lib.dptr(mx), _lib.dptr(my), _lib.dptr(mz), ny, nx, 0, _lib.dptr(x), _lib.dptr(y),
                                  _lib.dptr(dem.reshape(-1)), 2.0, 0, h))
    d_data = torch.from_numpy(data).cuda()
    d_out = torch.empty((T, ny * nx), dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    for det in (1, 0):
        for _ in range(2):
            _lib.check(L.smrfb_idw_distribute_dev(h, d_data.data_ptr(), T, det, -1, -73.0, 47.0, None, None, d_out.data_ptr(), st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            _lib.check(L.smrfb_idw_distribute_dev(h, d_data.data_ptr(), T, det, -1, -73.0, 47.0, None, None, d_out.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cs = ny * nx * T / (ms * 1e-3)
        print("IDW S=%d %dx%d T=%d detrend=%d: %.2f ms  %.2f G cell-steps/s  write %.1f GB/s  fp64 %.2f TFLOP/s" %
              (S, ny, nx, T, det, ms, cs / 1e9, cs * 8 / 1e9, cs * 2 * S / 1e12))
    print("finite:", bool(torch.isfinite(d_out).all()), float(d_out.mean()))
    if os.environ.get("IDW_COMPARE"):      # same call on the other kernel: max difference relative to the step scale
        os.environ["SMRFB_IDW_KERNEL"] = os.environ["IDW_COMPARE"]
        h2 = _lib.handle_t()
        _lib.check(L.smrfb_idw_create(S, _lib.dptr(mx), _lib.dptr(my), _lib.dptr(mz), ny, nx, 0, _lib.dptr(x), _lib.dptr(y),
                                      _lib.dptr(dem.reshape(-1)), 2.0, 0, h2))
        d_out2 = torch.empty_like(d_out)
        _lib.check(L.smrfb_idw_distribute_dev(h2, d_data.data_ptr(), T, 0, -1, -73.0, 47.0, None, None, d_out2.data_ptr(), st))
        torch.cuda.synchronize()
        scale = torch.from_numpy(np.nanmax(np.abs(data), axis=1)).cuda()[:, None]
        err = ((d_out - d_out2).abs() / scale)
        print("vs %s kernel: max rel diff %.3e  mean %.3e  nan mismatch %d" % (
            os.environ["IDW_COMPARE"], float(err.nan_to_num(0).max()), float(err.nan_to_num(0).mean()),
            int((torch.isnan(d_out) != torch.isnan(d_out2)).sum())))


if __name__ == "__main__":
    main()
