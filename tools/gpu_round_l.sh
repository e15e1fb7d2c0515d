#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 0 3 4; do
  UAV_IGEMM_B_STAGES=$n timeout 60 python tools/bench_linear512.py > gpurun_out/r2l_linear_bstages$n.txt 2>&1; cat gpurun_out/r2l_linear_bstages$n.txt
done
UAV_IGEMM_B_STAGES=3 timeout 100 python -m pytest tests/test_igemm_gpu.py -x -q -m gpu 2>&1 | tail -2
