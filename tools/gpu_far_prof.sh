#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_idw_far_gpu.py tests/test_idw_gpu.py -q > gpurun_out/far_pytest.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/far_pytest.log
DROP_MODE=iid IDW_COMPARE=dmma timeout 300 python tools/idw_far_perf.py 1024 10000 200 720 2>&1 | grep -E "info|IDW S|vs |finite|rror"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/far_launches.csv python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/far_ncu1.log 2>&1
grep -E "idw_far" gpurun_out/far_launches.csv | tail -3 | sed 's/.*smrfb::/smrfb::/' | cut -c1-60,200-
grep -E "idw_far" gpurun_out/far_launches.csv | tail -3 | awk -F'","' '{print $5, $NF}' | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/far_main python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/far_ncu2.log 2>&1
ls -la gpurun_out/far_main.ncu-rep
