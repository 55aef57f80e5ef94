#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_idw_far_gpu.py tests/test_idw_gpu.py tests/test_c4_domain_gpu.py tests/test_edges_gpu.py -m gpu -q -x ) > gpurun_out/far2_pytest.log 2>&1
tail -4 gpurun_out/far2_pytest.log
for v in "" "SMRFB_IDW_FAR_NN10=1"; do
  env $v timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', d['value']/1e9, d['per_field_ms'], d['parity_sample']['per_field']['precip'], d['roofline']['node_kernel_ms_per_launch'], d['roofline']['ms_per_launch'])"
done
