#!/bin/bash
# ncu evidence (bounded): launch list of one bench step (eager launches, cuDNN autotune off so the list is the step itself)
# and one full capture per hand-written hot kernel.  Every ncu command runs only after the same command exited 0 without
# ncu, and under `timeout`.
mkdir -p gpurun_out
BENCH="env DF_CUDNN_BENCHMARK=0 python bench.py --precision 3xtf32 --steps 1 --warmup 3 --frames 4 --no-graph --no-cpu-baseline --no-extras"
timeout 300 $BENCH > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python scripts/summarize_launches.py gpurun_out/launches.csv 5 > gpurun_out/launch_list_step.json; head -c 1500 gpurun_out/launch_list_step.json
timeout 200 python scripts/prof_kernels.py 3 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_q_kernel|loss_forward_kernel|knn1_d3_kernel' \
    -c 8 -o gpurun_out/prof_kernels python scripts/prof_kernels.py 1 > gpurun_out/ncu_prof.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
