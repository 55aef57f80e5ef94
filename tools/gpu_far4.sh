#!/bin/bash
# DFMA issue probe; C5 IDW with more near-station weights in shared memory (default) against the old 2 (SMRFB_FAR_NQM=2)
mkdir -p gpurun_out
./build/dfma_issue_probe | tee gpurun_out/dfma_issue_probe.txt
timeout 600 python -m pytest tests/test_idw_far_gpu.py tests/test_c4_domain_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in "SMRFB_FAR_NQM=2" "SMRFB_X=0"; do
  echo "== $v"
  env $v timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value']/1e9, d['per_field_ms'], d['parity_sample']['max_err'])"
  env $v timeout 600 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "info|IDW S"
done 2>&1 | tee gpurun_out/far4_variants.txt
