#!/bin/bash
# GRID grid_local: staged fold kernel against the direct one: parity tests, kernel timings, the C5 line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_edges_gpu.py tests/test_image_data_gpu.py -m gpu -q -x ) 2>&1 | tail -3
{
timeout 600 python tools/grid_perf.py 4000 4000 256 2>&1 | grep "GRID local"
SMRFB_GRID_FOLD_DIRECT=1 timeout 600 python tools/grid_perf.py 4000 4000 256 2>&1 | grep "GRID local"
for v in "SMRFB_X=1" "SMRFB_GRID_FOLD_DIRECT=1"; do
env $v timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', d['value']/1e9, d['per_field_ms'], d['parity_sample']['per_field']['vapor_pressure'], d['roofline']['per_field']['vapor_pressure']['frac'])"
done
} 2>&1 | tee gpurun_out/grid2_staged.txt
