#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py tests/test_conv_io_gpu.py -m gpu -q > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2e_tests.log | cut -c1-300
timeout 200 python tools/profile_unet.py > gpurun_out/r2e_unet_by_shape.txt 2>&1; grep -E "^\[|act2" gpurun_out/r2e_unet_by_shape.txt
UAV_IGEMM_DOUBLE_STAGING=0 timeout 200 python tools/profile_unet.py > gpurun_out/r2e_unet_by_shape_single_staging.txt 2>&1; grep -E "^\[|act2" gpurun_out/r2e_unet_by_shape_single_staging.txt
