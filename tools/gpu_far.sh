#!/bin/bash
# far-field IDW: parity tests, then timing on a C4-density window against the dense fp64 kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_idw_far_gpu.py tests/test_idw_gpu.py -x -q -s 2>&1 | tail -25 ) > gpurun_out/far_pytest.log 2>&1
tail -25 gpurun_out/far_pytest.log
for mode in iid block; do
  echo "== $mode"
  DROP_MODE=$mode IDW_COMPARE=dmma timeout 300 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "info|IDW S|vs |finite|rror" | tee -a gpurun_out/far_perf.txt
done
