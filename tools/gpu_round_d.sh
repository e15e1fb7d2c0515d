#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py -m gpu -q -k "layernorm_folded or concat_that or residual_tile" > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2d_tests.log | cut -c1-300
bash tools/gpu_round_mg1.sh
