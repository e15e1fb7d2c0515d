#!/bin/bash
# single-GPU baselines of the multi-GPU configs (denominators of the strong-scaling efficiencies)
set -u
mkdir -p gpurun_out
for c in c4 c3 clip64; do
  timeout 900 python bench.py --gpus 1 --config $c --steps 1 --warmup 1 --warmup-ddim-steps 2 --no-e2e > gpurun_out/r2_mg_${c}_n1.json 2> gpurun_out/r2_mg_${c}_n1.err
  echo "$c rc=$?"; cut -c1-700 gpurun_out/r2_mg_${c}_n1.json; tail -2 gpurun_out/r2_mg_${c}_n1.err
done
