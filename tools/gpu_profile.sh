#!/bin/bash
# ncu evidence for the bench: launch list of the bench command, full capture of the IDW kernel at the C4 slab shape
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s2_ncu_bench.log 2>&1
tail -2 gpurun_out/s2_ncu_bench.log | cut -c1-300
SMRFB_IDW_KERNEL=i8 timeout 600 ncu --set full --import-source on --clock-control none -k regex:idw_i8_kernel -c 1 -s 2 -f -o gpurun_out/s2_i8_c4slab \
    python tools/idw_perf.py 1000 10000 200 720 > gpurun_out/s2_ncu_i8_c4slab.log 2>&1
tail -3 gpurun_out/s2_ncu_i8_c4slab.log
ls -la gpurun_out/s2_launches.csv gpurun_out/s2_i8_c4slab.ncu-rep
