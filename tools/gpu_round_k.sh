#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -o gpurun_out/r2k_prof_linear_plain -f python tools/prof_kernels.py linear_plain > gpurun_out/r2k_prof_linear_plain.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r2k_prof_linear_plain.ncu-rep
