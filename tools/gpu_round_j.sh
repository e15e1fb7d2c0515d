#!/bin/bash
set -u
mkdir -p gpurun_out
for a in 0 8 16; do
  UAV_IGEMM_L2_AHEAD=$a timeout 120 python tools/bench_linear512.py > gpurun_out/r2j_linear_l2ahead$a.txt 2>&1; cat gpurun_out/r2j_linear_l2ahead$a.txt
done
UAV_IGEMM_L2_AHEAD=8 timeout 300 python -m pytest tests/test_igemm_gpu.py -x -q -m gpu 2>&1 | tail -3
