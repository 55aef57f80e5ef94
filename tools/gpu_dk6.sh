#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_distribute_mma_kernel -s 2 -c 1 -f -o gpurun_out/r02_dk_mma \
    python tools/dk_steady_perf.py > gpurun_out/r02_ncu_dk_mma.log 2>&1
tail -2 gpurun_out/r02_ncu_dk_mma.log
