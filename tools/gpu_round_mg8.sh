#!/bin/bash
# BASELINE configs 4 and 5 and the 64-frame strong-scaling clip on 8 GPUs (one process per GPU, NCCL)
set -u
mkdir -p gpurun_out
run() { # name, config
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --config $2 --steps 1 --warmup 1 --warmup-ddim-steps 2 --no-e2e > gpurun_out/r2_mg_$1_n8.json 2> gpurun_out/r2_mg_$1_n8.err
  echo "$1 rc=$?"; cut -c1-900 gpurun_out/r2_mg_$1_n8.json; tail -2 gpurun_out/r2_mg_$1_n8.err
}
run c4 c4
run clip64 clip64
run c5 c5
