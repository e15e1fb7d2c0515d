#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2f2_bench.json; tail -3 gpurun_out/r2f2_bench.err
timeout 200 python tools/profile_unet.py > gpurun_out/r2f2_unet_by_shape.txt 2>&1; grep "^\[" gpurun_out/r2f2_unet_by_shape.txt
timeout 300 python tools/profile_pipeline.py > gpurun_out/r2f2_pipeline_profile.txt 2>&1; head -12 gpurun_out/r2f2_pipeline_profile.txt
