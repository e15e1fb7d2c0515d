#!/bin/bash
# first GPU run of the next-rows draft: new kernels one by one, then the models, then timing.  Run under gpurun on one GPU.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_raft_gpu.py tests/test_clip_gpu.py -m gpu -q -x -s > gpurun_out/wip_tests.log 2>&1
tail -15 gpurun_out/wip_tests.log
timeout 300 python tools/bench_raft.py > gpurun_out/bench_raft.json 2> gpurun_out/bench_raft.err
cat gpurun_out/bench_raft.json; tail -3 gpurun_out/bench_raft.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/wip_full_gpu_tests.log 2>&1; tail -3 gpurun_out/wip_full_gpu_tests.log
