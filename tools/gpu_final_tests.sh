#!/bin/bash
# the whole GPU suite, smoke(), the DK build on the whole C4 grid
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
DK_PERF_REBUILDS=0 timeout 900 python tools/dk_perf.py 10000 10000 200 0 0 2>&1 | grep "smrfb\|build" | tee gpurun_out/r02_dk_build_whole_grid.txt
