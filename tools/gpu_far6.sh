#!/bin/bash
# far kernel after an instruction diet: parity tests, timing at C4 and C5 station density, executed warp instructions
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_idw_far_gpu.py tests/test_c4_domain_gpu.py tests/test_edges_gpu.py -m gpu -q -x 2>&1 | tail -2
{
timeout 600 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "IDW S"
timeout 600 python tools/idw_far_perf.py 4000 4000 200 512 10 4000 2>&1 | grep -E "IDW S"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum -k regex:idw_far_kernel -s 2 -c 1 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "inst_executed|duration"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum -k regex:idw_far_kernel -s 2 -c 1 python tools/idw_far_perf.py 4000 4000 200 512 10 4000 2>&1 | grep -E "inst_executed|duration"
} 2>&1 | tee gpurun_out/far6_${FAR_TAG:-x}.txt
