"""Device-resident timing of the IDW kernels on a WINDOW of the C4 mesh (30 m cells inside a 300 km network of 200
stations, i.e. the station density of BASELINE.json configs[3]) -- development aid, not the bench.
usage: idw_far_perf.py [ny nx S T dx domain_cells]   (C5 density: 4000 4000 200 512 10 4000)   env: SMRFB_IDW_KERNEL (unset = default), DROP_MODE=block|iid, IDW_COMPARE=dmma|i8"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from smrf_b200 import _lib, synthetic as syn  # noqa: E402


def main():
    ny = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 720
    rng = np.random.default_rng(7)
    dx = float(sys.argv[5]) if len(sys.argv) > 5 else 30.0
    dom = int(sys.argv[6]) if len(sys.argv) > 6 else 10000
    x = 500000.0 + dx * np.arange(nx)
    y = 4800000.0 - (97000.0 if dom == 10000 else 0.0) - dx * np.arange(ny)
    X, Y = x[None, :], y[:, None]
    dem = (1800.0 + 700.0 * np.sin(2 * np.pi * X / 41000.0) * np.cos(2 * np.pi * Y / 57000.0)
           + 120.0 * np.sin(2 * np.pi * X / 3100.0 + 1.0) * np.sin(2 * np.pi * Y / 2300.0)).astype(np.float32).astype(np.float64)
    mx = 500000.0 + dx * rng.integers(0, dom, S) + 0.37 * dx
    my = 4800000.0 - dx * rng.integers(0, dom, S) + 0.37 * dx
    mz = 1800.0 + 700.0 * np.sin(2 * np.pi * mx / 41000.0) * np.cos(2 * np.pi * my / 57000.0) + rng.normal(0, 5.0, S)
    if os.environ.get("DROP_MODE") == "block":
        data, nm = syn.dropouts_block(syn.make_series("air_temp", mz, T), max_masks=12, per_year=40.0)
    else:
        data = syn.dropouts_iid(syn.make_series("air_temp", mz, T), p=0.02)
    L = _lib.lib()
    h = _lib.handle_t()
    _lib.check(L.smrfb_idw_create(S, _lib.dptr(mx), _lib.dptr(my), _lib.dptr(mz), ny, nx, 0, _lib.dptr(x), _lib.dptr(y),
                                  _lib.dptr(dem.reshape(-1)), 2.0, 0, h))
    info = np.zeros(8, dtype=np.int32)
    _lib.check(L.smrfb_idw_info(h, info.ctypes.data_as(_lib.c_i32_p)))
    print("info: far %d nodes %d max_near %d patches %d theta %.1f i8 %d" % (info[0], info[1], info[2], info[3], info[4] / 1e3, info[5]))
    d_data = torch.from_numpy(data).cuda()
    d_out = torch.empty((T, ny * nx), dtype=torch.float64, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    for det in (1,):
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
        print("IDW S=%d %dx%d T=%d detrend=%d: %.2f ms  %.2f G cell-steps/s  write %.1f GB/s = %.3f of 6538" %
              (S, ny, nx, T, det, ms, cs / 1e9, cs * 8 / 1e9, cs * 8 / 1e9 / 6538.3))
    print("finite:", bool(torch.isfinite(d_out).all()), float(d_out.mean()))
    cmp = os.environ.get("IDW_COMPARE")
    if cmp:
        Tc = min(T, 48)
        os.environ["SMRFB_IDW_KERNEL"] = cmp
        h2 = _lib.handle_t()
        _lib.check(L.smrfb_idw_create(S, _lib.dptr(mx), _lib.dptr(my), _lib.dptr(mz), ny, nx, 0, _lib.dptr(x), _lib.dptr(y),
                                      _lib.dptr(dem.reshape(-1)), 2.0, 0, h2))
        d_a = d_out[:Tc]
        _lib.check(L.smrfb_idw_distribute_dev(h, d_data.data_ptr(), Tc, 0, -1, -73.0, 47.0, None, None, d_a.data_ptr(), st))
        d_b = torch.empty_like(d_a)
        _lib.check(L.smrfb_idw_distribute_dev(h2, d_data.data_ptr(), Tc, 0, -1, -73.0, 47.0, None, None, d_b.data_ptr(), st))
        torch.cuda.synchronize()
        scale = torch.from_numpy(np.nanmax(np.abs(data[:Tc]), axis=1)).cuda()[:, None]
        err = ((d_a - d_b).abs() / scale)
        print("vs %s kernel (%d steps): max rel diff %.3e  mean %.3e  nan mismatch %d" % (
            cmp, Tc, float(err.nan_to_num(0).max()), float(err.nan_to_num(0).mean()),
            int((torch.isnan(d_a) != torch.isnan(d_b)).sum())))


if __name__ == "__main__":
    main()
