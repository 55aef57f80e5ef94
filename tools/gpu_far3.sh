#!/bin/bash
# far-field kernel variants (CTAs per SM, software-pipelined phase B) + the snow consumer tests
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_consumers_gpu.py -m gpu -q -x ) > gpurun_out/snow_pytest.log 2>&1
tail -4 gpurun_out/snow_pytest.log
for v in "SMRFB_FAR_MINB=2 SMRFB_FAR_PIPE=0" "SMRFB_FAR_MINB=2 SMRFB_FAR_PIPE=1" "SMRFB_FAR_MINB=3 SMRFB_FAR_PIPE=0" "SMRFB_FAR_MINB=3 SMRFB_FAR_PIPE=1"; do
  echo "== $v"
  env $v timeout 600 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "IDW S|finite"
  env $v timeout 600 python -m pytest tests/test_idw_far_gpu.py -m gpu -q -x 2>&1 | tail -1
done 2>&1 | tee gpurun_out/far3_variants.txt
