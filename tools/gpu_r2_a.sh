#!/bin/bash
# Round 2, GPU visit A: parity tests (new: cubic, C4-domain int8 IDW, advisor regressions), DK build timing, GRID timings, D2H probe.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -15 gpurun_out/r2a_pytest.log
( timeout 600 python tools/dk_perf.py 1000 1000 200 ) > gpurun_out/r2a_dk_perf.txt 2>&1
tail -8 gpurun_out/r2a_dk_perf.txt
( timeout 600 python tools/grid_perf.py 4000 4000 256 ) > gpurun_out/r2a_grid_perf.txt 2>&1
tail -12 gpurun_out/r2a_grid_perf.txt
( timeout 120 python tools/d2h_probe.py ) > gpurun_out/r2a_d2h_n1.json 2> gpurun_out/r2a_d2h_n1.err
cat gpurun_out/r2a_d2h_n1.json; tail -3 gpurun_out/r2a_d2h_n1.err
free -g | head -2; nproc; lscpu | grep -i "numa\|model name\|socket" | head -8
