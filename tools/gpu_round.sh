#!/bin/bash
# One GPU-box visit: parity tests, probes, kernel timings, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2_pytest.log 2>&1
tail -3 gpurun_out/s2_pytest.log
timeout 120 ./build/umma_i8_probe > gpurun_out/s2_umma_i8_probe.txt 2>&1
for mode in block iid; do
  echo "== $mode" >> gpurun_out/s2_idw_perf.txt
  DROP_MODE=$mode SMRFB_IDW_KERNEL=i8 IDW_COMPARE=dmma timeout 300 python tools/idw_perf.py 2000 2000 200 720 >> gpurun_out/s2_idw_perf.txt 2>&1
done
cat gpurun_out/s2_idw_perf.txt
( time timeout 900 python bench.py ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
tail -c 1500 gpurun_out/s2_bench.json; tail -5 gpurun_out/s2_bench.err
