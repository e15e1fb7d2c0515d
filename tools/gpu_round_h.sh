#!/bin/bash
set -u
mkdir -p gpurun_out
UAV_GN_SILU_PAIR=0 timeout 200 python tools/bench_gn.py > gpurun_out/r2h_gn_single.txt 2>&1; cat gpurun_out/r2h_gn_single.txt
timeout 200 python tools/bench_gn.py > gpurun_out/r2h_gn_pair.txt 2>&1; cat gpurun_out/r2h_gn_pair.txt
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_igemm_gpu.py -x -q -m gpu -k "norm or gn or group" 2>&1 | tail -5
timeout 200 python tools/profile_unet.py > gpurun_out/r2h_unet_by_shape.txt 2>&1; grep "^\[" gpurun_out/r2h_unet_by_shape.txt
