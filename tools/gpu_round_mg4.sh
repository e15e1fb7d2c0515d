#!/bin/bash
# BASELINE config 3 (32 frames, 4 GPUs)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --config c3 --steps 1 --warmup 1 --warmup-ddim-steps 2 --no-e2e > gpurun_out/r2_mg_c3_n4.json 2> gpurun_out/r2_mg_c3_n4.err
echo "c3 rc=$?"; cut -c1-900 gpurun_out/r2_mg_c3_n4.json; tail -2 gpurun_out/r2_mg_c3_n4.err
