#!/bin/bash
# ncu evidence for profiles/: three `--set full` captures (GEMM as conv and as Linear, attention kernels) and the launch list of
# a 2-DDIM-step bench run.  Run under gpurun on ONE GPU; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -o gpurun_out/prof_conv_2sm -f python tools/prof_conv.py > gpurun_out/prof_conv_2sm.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -o gpurun_out/prof_linear_2sm -f python tools/prof_linear.py > gpurun_out/prof_linear_2sm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:fa_tc|temporal' -s 2 -c 2 -o gpurun_out/prof_attn -f python tools/prof_attn.py > gpurun_out/prof_attn.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline > gpurun_out/ncu_bench_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_v2.csv
tail -2 gpurun_out/prof_attn.log
