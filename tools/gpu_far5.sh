#!/bin/bash
# C5-density IDW: pipelined or not, then a full ncu capture of the far kernel there
mkdir -p gpurun_out
for v in "SMRFB_FAR_PIPE=0" "SMRFB_FAR_PIPE=1"; do
  echo "== $v"
  env $v timeout 600 python tools/idw_far_perf.py 4000 4000 200 512 10 4000 2>&1 | grep -E "info|IDW S"
done 2>&1 | tee gpurun_out/far5_variants.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/r02_far_c5 \
    python tools/idw_far_perf.py 4000 4000 200 512 10 4000 > gpurun_out/r02_ncu_far_c5.log 2>&1
tail -3 gpurun_out/r02_ncu_far_c5.log
