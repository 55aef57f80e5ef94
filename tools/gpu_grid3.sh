#!/bin/bash
# GRID cubic: staged kernel (two cells / one cell per thread) against the direct one
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_edges_gpu.py tests/test_image_data_gpu.py tests/test_consumers_gpu.py -m gpu -q -x ) 2>&1 | tail -3
{
timeout 900 python tools/grid_perf.py 4000 4000 256 2>&1 | grep -i "cubic"
SMRFB_GRID_CUBIC_CPT1=1 timeout 900 python tools/grid_perf.py 4000 4000 256 2>&1 | grep -i "cubic"
SMRFB_GRID_CUBIC_DIRECT=1 timeout 900 python tools/grid_perf.py 4000 4000 256 2>&1 | grep -i "cubic"
} 2>&1 | tee gpurun_out/grid3_cubic.txt
