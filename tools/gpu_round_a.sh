#!/bin/bash
# round-2 GPU call A: whole GPU suite (incl. the new RAFT / CLIP kernels and the full-size parity tests), bench, reference arm
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2a_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2a_gpu_tests.log | tail -3
grep -E "^\[|FAILED|Error" gpurun_out/r2a_gpu_tests.log | head -60
timeout 300 python tools/bench_raft.py > gpurun_out/r2a_bench_raft.json 2> gpurun_out/r2a_bench_raft.err
cat gpurun_out/r2a_bench_raft.json; tail -3 gpurun_out/r2a_bench_raft.err
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/r2a_bench_ref.json
