#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_consumers_gpu.py tests/test_image_data_gpu.py -x -q > gpurun_out/cons_pytest.log 2>&1; tail -30 gpurun_out/cons_pytest.log
