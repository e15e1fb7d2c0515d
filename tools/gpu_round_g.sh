#!/bin/bash
# FlashAttention softmax rewrite: parity (small + BASELINE size), timing, ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" > gpurun_out/r2g_tests.log 2>&1
echo "attention tests rc=$?"; tail -4 gpurun_out/r2g_tests.log | cut -c1-300
timeout 600 python -m pytest "tests/test_fullsize_gpu.py::test_vae3d_decode_chunk_config2" "tests/test_fullsize_gpu.py::test_unet_forward_config2" tests/test_vae_pipeline_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed" | cut -c1-400
timeout 300 python tools/bench_attention.py 2>&1 | tail -6
for k in fa512 fa128; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fa_tc -s 2 -c 1 -o gpurun_out/r2g_prof_$k -f python tools/prof_kernels.py $k > gpurun_out/r2g_prof_$k.log 2>&1; echo "$k rc=$?"
done
