#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_idw_far_gpu.py -x -q 2>&1 | tail -5 ) 
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/far_launches.csv python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/far_ncu1.log 2>&1
grep -E "idw_far|prep" gpurun_out/far_launches.csv | tail -8
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_kernel -s 2 -c 1 -f -o gpurun_out/far_main python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/far_ncu2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:idw_far_node_kernel -s 2 -c 1 -f -o gpurun_out/far_node python tools/idw_far_perf.py 1024 10000 200 720 > gpurun_out/far_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
