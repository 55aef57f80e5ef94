#!/bin/bash
# the bench lines and launch lists of tools/gpu_r2_final.sh (re-run after a bench.py fix)
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -3 gpurun_out/r02_bench_c4.err; cut -c1-250 gpurun_out/r02_bench_c4.json
( time timeout 900 python bench.py --workload c5 ) > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -3 gpurun_out/r02_bench_c5.err; cut -c1-250 gpurun_out/r02_bench_c5.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-300 gpurun_out/r02_bench_ref.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c5.log 2>&1
ls -la gpurun_out/r02_bench* gpurun_out/r02_launches*
