#!/bin/bash
# DK weight build: chain walk: parity tests, then an ncu source-level capture on a small region.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dk_gpu.py tests/test_edges_gpu.py tests/test_c4_domain_gpu.py -m gpu -q ) > gpurun_out/dk2_pytest.log 2>&1
tail -5 gpurun_out/dk2_pytest.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dk_chain_kernel -c 1 -f -o gpurun_out/dk2_chain \
    python tools/dk_perf.py 64 1024 200 > gpurun_out/dk2_ncu.log 2>&1
tail -3 gpurun_out/dk2_ncu.log
