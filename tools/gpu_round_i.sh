#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/bench_linear512.py > gpurun_out/r2i_linear_default.txt 2>&1; cat gpurun_out/r2i_linear_default.txt
UAV_IGEMM_RES_MODE=1 timeout 120 python tools/bench_linear512.py > gpurun_out/r2i_linear_res1.txt 2>&1; grep "residual=1\|UAV" gpurun_out/r2i_linear_res1.txt
UAV_IGEMM_A_PROMO=128 timeout 120 python tools/bench_linear512.py > gpurun_out/r2i_linear_promo128.txt 2>&1; cat gpurun_out/r2i_linear_promo128.txt
UAV_IGEMM_DEEP_A=1 timeout 120 python tools/bench_linear512.py > gpurun_out/r2i_linear_deepa.txt 2>&1; cat gpurun_out/r2i_linear_deepa.txt
