#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:grid_local4_kernel -s 1 -c 1 -f -o gpurun_out/r02_grid_local4 \
    python tools/grid_perf.py 4000 4000 256 > gpurun_out/r02_ncu_grid_local4.log 2>&1
tail -2 gpurun_out/r02_ncu_grid_local4.log
