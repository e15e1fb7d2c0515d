#!/bin/bash
# round-2 GPU call C: virtual concat, fused UNet tail (+ sampler epilogue), fence-free accumulator hand-back
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_igemm_gpu.py tests/test_conv_io_gpu.py tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s > gpurun_out/r2c_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2c_gpu_tests.log | tail -3
grep -E "^\[|FAILED|Error|^E  " gpurun_out/r2c_gpu_tests.log | head -60
timeout 400 python tools/profile_unet.py --ab > gpurun_out/r2c_unet_by_shape.txt 2>&1; grep "^\[" gpurun_out/r2c_unet_by_shape.txt
timeout 600 python bench.py --steps 2 --warmup 2 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; cut -c1-1800 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
