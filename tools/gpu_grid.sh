#!/bin/bash
# GRID: parity tests, then the kernel timings (tools/grid_perf.py: 600 m source spacing; bench c5: 3 km)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_edges_gpu.py tests/test_image_data_gpu.py tests/test_consumers_gpu.py -m gpu -q -x ) > gpurun_out/grid_pytest.log 2>&1
tail -4 gpurun_out/grid_pytest.log
( timeout 600 python tools/grid_perf.py 4000 4000 256 ) 2>&1 | grep "GRID local"
GRID_LOCAL_EXACT=1 timeout 600 python tools/grid_perf.py 4000 4000 256 2>&1 | grep "GRID local"
timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value']/1e9, d['per_field_ms'], d['parity_sample']['per_field']['vapor_pressure'], d['roofline']['per_field']['vapor_pressure']['frac'])"
