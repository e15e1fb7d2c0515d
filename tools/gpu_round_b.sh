#!/bin/bash
# round-2 GPU call B: kernel changes of the round (GN statistics from the producers, in-place concat, TMA residual, GELU,
# VAE stream scale) — tests, per-shape UNet profile with A/B switches, bench, ncu captures
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2b_gpu_tests.log | tail -3
grep -E "^\[|FAILED|Error|assert" gpurun_out/r2b_gpu_tests.log | head -70
timeout 300 python tools/profile_unet.py --ab > gpurun_out/r2b_unet_by_shape.txt 2>&1; grep "^\[" gpurun_out/r2b_unet_by_shape.txt
UAV_IGEMM_RES_MODE=1 timeout 200 python tools/profile_unet.py > gpurun_out/r2b_unet_by_shape_resmode1.txt 2>&1; grep "^\[" gpurun_out/r2b_unet_by_shape_resmode1.txt
timeout 600 python bench.py --steps 2 --warmup 2 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; cat gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
bash tools/run_ncu_round2.sh geglu linear_res conv_gn gn_fused fa512 fa128 cross
