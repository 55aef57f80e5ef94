#!/bin/bash
# GRID: parity tests, then the kernel timings on the C5 shape (tools/grid_perf.py: 600 m source spacing; bench c5: 3 km)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_edges_gpu.py tests/test_image_data_gpu.py -m gpu -q -x ) > gpurun_out/grid_pytest.log 2>&1
tail -4 gpurun_out/grid_pytest.log
( timeout 600 python tools/grid_perf.py 4000 4000 256 ) > gpurun_out/grid_perf.txt 2>&1
grep "GRID local" gpurun_out/grid_perf.txt
SMRFB_GRID_TWO_CELLS=1 timeout 600 python tools/grid_perf.py 4000 4000 256 2>&1 | grep "GRID local"
for v in "" "SMRFB_GRID_TWO_CELLS=1"; do
  env $v timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', d['value']/1e9, d['per_field_ms'], d['parity_sample']['max_err'])"
done
