#!/bin/bash
# ncu evidence (launch list + one full capture per hot kernel) and encoder-variant timings.
mkdir -p gpurun_out
timeout 600 python scripts/encoder_variants.py > gpurun_out/encoder_variants.log 2>&1; tail -n 6 gpurun_out/encoder_variants.log
BENCH="env DF_CUDNN_BENCHMARK=0 python bench.py --precision 3xtf32 --steps 2 --warmup 3 --frames 4 --no-graph --no-cpu-baseline --no-extras"
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python scripts/prof_kernels.py 3 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|loss_forward_kernel|knn1_d3_kernel|sgemm_kernel' \
    -c 40 -o gpurun_out/prof_kernels python scripts/prof_kernels.py 1 > gpurun_out/ncu_prof.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
