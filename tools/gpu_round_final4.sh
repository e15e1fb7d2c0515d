#!/bin/bash
# last validation of the round on one GPU: whole GPU suite, smoke, short bench (no reference legs)
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu -s -x > gpurun_out/r2f4_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/r2f4_gpu_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 2 --warmup 3 --no-reference-gpu --no-cpu-baseline > gpurun_out/r2f4_bench.json 2> gpurun_out/r2f4_bench.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/r2f4_bench.json
