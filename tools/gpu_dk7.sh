#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_dk_gpu.py tests/test_consumers_gpu.py tests/test_c4_domain_gpu.py -m gpu -q -x 2>&1 | tail -3 )
DK_PERF_REBUILDS=1 timeout 600 python tools/dk_perf.py 512 1024 200 3000 4000 2>&1 | grep "smrfb" | tee gpurun_out/dk7_phases.txt
