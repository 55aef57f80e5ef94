#!/bin/bash
# Round 2 ncu evidence: launch lists of both bench workloads (same commands as the bench lines, one timed step), full
# captures of the dominant kernels at the bench's launch shapes.  Nothing printed under ncu is a bench value.
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c4.log 2>&1
tail -1 gpurun_out/r02_ncu_bench_c4.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c5.log 2>&1
tail -1 gpurun_out/r02_ncu_bench_c5.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/r02_far_c4slab \
    python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/r02_ncu_far.log 2>&1
tail -2 gpurun_out/r02_ncu_far.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_distribute_kernel -s 2 -c 1 -f -o gpurun_out/r02_dk_steady \
    python tools/dk_steady_perf.py > gpurun_out/r02_ncu_dk.log 2>&1
tail -2 gpurun_out/r02_ncu_dk.log
ls -la gpurun_out/r02_*
