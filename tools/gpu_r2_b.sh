#!/bin/bash
# Round 2, GPU visit B: remaining parity tests, GRID timings, ncu source-level capture of dk_build_kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
( timeout 600 python tools/grid_perf.py 4000 4000 256 ) > gpurun_out/r2b_grid_perf.txt 2>&1
tail -4 gpurun_out/r2b_grid_perf.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_build_kernel -c 1 -f -o gpurun_out/r2b_dk_build \
    python tools/dk_perf.py 300 1000 200 > gpurun_out/r2b_ncu_dk.log 2>&1
tail -3 gpurun_out/r2b_ncu_dk.log
ls -la gpurun_out/r2b_dk_build.ncu-rep
