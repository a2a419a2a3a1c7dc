#!/bin/bash
# ncu --set full capture (with source-level stall sampling) of one launch per case: gpu_prof_cases.sh <tag> case...
tag=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 100 python scripts/prof_case.py $c hybrid16s 2 > gpurun_out/${tag}_${c}_plain.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_q_kernel --launch-skip 1 --launch-count 1 \
      -o gpurun_out/${tag}_${c} python scripts/prof_case.py $c hybrid16s 2 > gpurun_out/${tag}_${c}_ncu.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/${tag}_*.ncu-rep
