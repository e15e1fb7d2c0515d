#!/bin/bash
# last validation of the round on one GPU: whole GPU suite, smoke, whole-clip profile, ncu launch list of one UNet forward
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/r2f3_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/r2f3_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 300 python tools/profile_pipeline.py > gpurun_out/r2f3_pipeline_profile.txt 2>&1; head -12 gpurun_out/r2f3_pipeline_profile.txt
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_unet_forward.csv python tools/ncu_unet_forward.py > gpurun_out/r2f3_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2_launches_unet_forward.csv
