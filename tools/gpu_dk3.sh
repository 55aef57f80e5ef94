#!/bin/bash
# DK weight build over the WHOLE C4 grid (1e8 cells): chain walk, then the tree walk.
mkdir -p gpurun_out
for walk in chain tree; do
  echo "== $walk" >> gpurun_out/dk4_perf.txt
  DK_PERF_REBUILDS=0 SMRFB_DK_WALK=$walk timeout 900 python tools/dk_perf.py 10000 10000 200 0 0 2>&1 | grep -v "^8 steps" >> gpurun_out/dk4_perf.txt
done
cat gpurun_out/dk4_perf.txt
