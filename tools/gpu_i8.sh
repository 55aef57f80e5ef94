#!/bin/bash
# quick visit: i8 parity tests + timing of both dropout modes against the DMMA kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_idw_i8_gpu.py tests/test_idw_gpu.py -x -q 2>&1 | tail -3
for mode in block iid; do
  echo "== $mode"
  DROP_MODE=$mode SMRFB_IDW_KERNEL=i8 IDW_COMPARE=dmma timeout 300 python tools/idw_perf.py 2000 2000 200 720 2>&1 | grep -E "IDW S|vs |prof|rror"
done
