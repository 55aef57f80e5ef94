#!/bin/bash
# full-size C4 bench line (device-timed part only) with the DK build statistics on stderr
mkdir -p gpurun_out
( time SMRFB_DK_VERBOSE=1 timeout 1500 python bench.py --no-cpu-baseline --no-e2e --no-extras ) > gpurun_out/bench_c4_dev.json 2> gpurun_out/bench_c4_dev.err
tail -8 gpurun_out/bench_c4_dev.err; cut -c1-300 gpurun_out/bench_c4_dev.json
