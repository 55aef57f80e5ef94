#!/bin/bash
# far-field node kernel on the fp64 tensor cores against the DFMA version: parity tests, field time at C4 and C5 station density
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_idw_far_gpu.py tests/test_c4_domain_gpu.py tests/test_edges_gpu.py tests/test_idw_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in "SMRFB_FAR_NODE=fma" "SMRFB_FAR_NODE=mma"; do
  echo "== $v"
  env $v timeout 600 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "IDW S"
  env $v timeout 600 python tools/idw_far_perf.py 4000 4000 200 512 10 4000 2>&1 | grep -E "IDW S"
done 2>&1 | tee gpurun_out/far7_node.txt
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value']/1e9, d['per_field_ms'], d['parity_sample']['max_err'], d['roofline']['ms_per_launch'], d['roofline']['node_kernel_ms_per_launch'])" | tee -a gpurun_out/far7_node.txt
