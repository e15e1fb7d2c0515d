#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "n2 rc=$?"; cut -c1-1500 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
