#!/bin/bash
# DK weight build: parity tests, timings on windows of the C4 basin.  SMRFB_DK_WALK = sets | chain | tree
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_dk_gpu.py tests/test_edges_gpu.py tests/test_c4_domain_gpu.py -m gpu -q -x ) > gpurun_out/dk2_pytest.log 2>&1
tail -5 gpurun_out/dk2_pytest.log
rm -f gpurun_out/dk2_perf.txt
for reg in "512 1024 200 3000 4000" "2048 2048 200 3000 7900"; do
    echo "== $reg" >> gpurun_out/dk2_perf.txt
    DK_PERF_REBUILDS=1 timeout 600 python tools/dk_perf.py $reg 2>&1 | grep -v "^8 steps" >> gpurun_out/dk2_perf.txt
done
cat gpurun_out/dk2_perf.txt
