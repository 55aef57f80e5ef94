#!/bin/bash
# bench.py on the GPU box: a reduced-size pass of both workloads first (catches errors in seconds), then the full sizes
mkdir -p gpurun_out
echo "== reduced c4"; timeout 600 python bench.py --ny 2048 --nx 2048 --tb 64 --steps 2 --warmup 1 > gpurun_out/bench_small_c4.json 2> gpurun_out/bench_small_c4.err; tail -3 gpurun_out/bench_small_c4.err; cut -c1-600 gpurun_out/bench_small_c4.json
echo "== reduced c5"; timeout 600 python bench.py --workload c5 --ny 1000 --nx 1000 --tb 32 --steps 2 --warmup 1 > gpurun_out/bench_small_c5.json 2> gpurun_out/bench_small_c5.err; tail -3 gpurun_out/bench_small_c5.err; cut -c1-600 gpurun_out/bench_small_c5.json
if [ "$1" != "small" ]; then
echo "== full c4"; ( time timeout 1200 python bench.py ) > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -4 gpurun_out/bench_c4.err; cut -c1-400 gpurun_out/bench_c4.json
echo "== full c5"; ( time timeout 900 python bench.py --workload c5 ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -4 gpurun_out/bench_c5.err; cut -c1-400 gpurun_out/bench_c5.json
fi
