#!/bin/bash
# N ranks on one box: what the host can take (D2H probe), then the bench line at N GPUs (torchrun, as the driver launches it)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
lscpu | grep -i "numa\|model name\|socket\|^CPU(s)" > gpurun_out/r02_lscpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py \
    > gpurun_out/r02_d2h_n$N.json 2> gpurun_out/r02_d2h_n$N.err
tail -c 600 gpurun_out/r02_d2h_n$N.json
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py \
    --gpus $N --steps 3 --warmup 3 --no-extras ) > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err
tail -5 gpurun_out/r02_bench_c4_n$N.err; tail -c 1500 gpurun_out/r02_bench_c4_n$N.json
