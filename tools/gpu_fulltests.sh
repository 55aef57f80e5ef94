#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/full_pytest.log 2>&1; tail -12 gpurun_out/full_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
