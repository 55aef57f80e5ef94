#!/bin/bash
# DK steady state: tensor-core kernel vs ELL kernel (parity tests, then timing at a C4 slab shape)
mkdir -p gpurun_out
if [ "$1" != "perf" ]; then
( time timeout 900 python -m pytest tests/test_dk_gpu.py tests/test_edges_gpu.py tests/test_c4_domain_gpu.py tests/test_image_data_gpu.py tests/test_consumers_gpu.py -m gpu -q -x ) > gpurun_out/dk5_pytest.log 2>&1
tail -5 gpurun_out/dk5_pytest.log
fi
for m in mma ell; do echo "== $m"; SMRFB_DK_STEADY=$m timeout 600 python tools/dk_steady_perf.py 2>&1 | tail -1; done | tee gpurun_out/dk5_perf.txt
