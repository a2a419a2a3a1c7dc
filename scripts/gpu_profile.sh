#!/bin/bash
# ncu evidence (bounded): launch list of ONE eager step of the bench workload (scripts/prof_step.py, between profiler marks)
# and one full capture per hand-written hot kernel.  Every ncu command runs only after the same command exited 0 without
# ncu, and under `timeout`.
mkdir -p gpurun_out
STEP="python scripts/prof_step.py 32 hybrid16s"
timeout 300 $STEP > gpurun_out/prof_step.txt 2>&1 &&
DF_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv $STEP > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_list_step.json; head -c 1500 gpurun_out/launch_list_step.json
timeout 200 python scripts/prof_kernels.py 3 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_q_kernel|upconv_finish_kernel|loss_forward_kernel|knn1_d3_kernel' \
    -c 11 -o gpurun_out/prof_kernels python scripts/prof_kernels.py 1 > gpurun_out/ncu_prof.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
