#!/bin/bash
# Round 2 final evidence: the whole GPU suite + smoke(), full-size bench lines, ncu launch lists of the same commands (one timed
# step), full captures of the dominant kernels at the bench's launch shapes.  Nothing printed under ncu is a bench value.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -3 gpurun_out/r02_bench_c4.err; cut -c1-250 gpurun_out/r02_bench_c4.json
( time timeout 900 python bench.py --workload c5 ) > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; tail -3 gpurun_out/r02_bench_c5.err; cut -c1-250 gpurun_out/r02_bench_c5.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-300 gpurun_out/r02_bench_ref.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench_c5.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/r02_far_c4slab \
    python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/r02_ncu_far.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/r02_far_c5 \
    python tools/idw_far_perf.py 4000 4000 200 512 10 4000 > gpurun_out/r02_ncu_far_c5.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:grid_local_fold_staged_kernel -s 1 -c 1 -f -o gpurun_out/r02_grid_fold \
    python tools/grid_perf.py 4000 4000 256 > gpurun_out/r02_ncu_grid_fold.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_node_mma_kernel -s 2 -c 1 -f -o gpurun_out/r02_far_node \
    python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/r02_ncu_far_node.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_distribute_mma_kernel -s 2 -c 1 -f -o gpurun_out/r02_dk_mma \
    python tools/dk_steady_perf.py > gpurun_out/r02_ncu_dk_mma.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_sets_kernel -c 1 -f -o gpurun_out/r02_dk_sets \
    python tools/dk_perf.py 64 1024 200 > gpurun_out/r02_ncu_dk_sets.log 2>&1
DK_PERF_REBUILDS=0 timeout 900 python tools/dk_perf.py 10000 10000 200 0 0 2>&1 | grep "smrfb\|build" | tee gpurun_out/r02_dk_build_whole_grid.txt
ls -la gpurun_out/r02_*
