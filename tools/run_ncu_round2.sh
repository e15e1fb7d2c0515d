#!/bin/bash
# round-2 ncu evidence: `--set full` captures (with source) of the kernels VERDICT r1 names, then the launch list of a
# 2-DDIM-step bench run.  ONE GPU; numbers printed under ncu are never bench values.  usage: run_ncu_round2.sh [kernels...]
set -u
mkdir -p gpurun_out
KS=${@:-"geglu linear_res conv_gn gn_fused fa512 fa128 cross"}
for k in $KS; do
  case $k in
    geglu|linear_res|linear_res_gn|conv_gn|conv) pat='regex:igemm' ;;
    gn_fused) pat='regex:gn_reduce|gn_apply' ;;
    gn_plain) pat='regex:gn_' ;;
    fa512|fa128) pat='regex:fa_tc' ;;
    cross) pat='regex:cross_attn' ;;
  esac
  n=1; [ $k = gn_fused ] && n=2; [ $k = gn_plain ] && n=3
  skip=$((2*n))
  timeout 300 ncu --set full --clock-control none --import-source on -k "$pat" -s $skip -c $n -o gpurun_out/r2_prof_$k -f \
      python tools/prof_kernels.py $k > gpurun_out/r2_prof_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/r2_prof_*.ncu-rep
