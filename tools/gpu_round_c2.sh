#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_igemm_gpu.py::test_groupnorm_of_a_concat_that_is_never_built tests/test_conv_io_gpu.py tests/test_vae_pipeline_gpu.py "tests/test_fullsize_gpu.py::test_pipeline_config2_short" -m gpu -q -s > gpurun_out/r2c2_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2c2_gpu_tests.log | tail -3
grep -E "^\[|FAILED|Error|^E  " gpurun_out/r2c2_gpu_tests.log | head -40
timeout 600 python bench.py --steps 2 --warmup 2 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; cut -c1-2600 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
