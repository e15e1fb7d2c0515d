#!/bin/bash
# round-2 final validation on one GPU: whole GPU suite, smoke, the driver's bench commands, ncu evidence
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2f_gpu_tests.log | tail -3; grep -E "FAILED|^E  " gpurun_out/r2f_gpu_tests.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
echo "reference arm rc=$?"; cut -c1-900 gpurun_out/r2f_bench_ref.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -o gpurun_out/r2_prof_conv -f python tools/prof_kernels.py conv > gpurun_out/r2_prof_conv.log 2>&1
echo "ncu conv rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_2step.csv python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline --no-reference-gpu --no-e2e > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"; ls -la gpurun_out/r2_launches_2step.csv gpurun_out/r2_prof_conv.ncu-rep
