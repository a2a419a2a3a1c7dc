#!/bin/bash
# ncu --set full of the non-GEMM kernels of one eager step (the plain run of the same command exited 0 in this session's CI):
# pyramid pool / sum, patch gather, upconv_finish, max pool, stride-2 patches, log-softmax, select_out -- one pass, first launch of each.
mkdir -p gpurun_out
DF_NCU=1 timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"pyramid|gather_up|upconv_finish|maxpool|im2col_s2|log_softmax|select_out|xyz_conv|pool_finish" --launch-count 24 \
   -o gpurun_out/r2s5_helpers python scripts/prof_step.py 32 hybrid16s > gpurun_out/r2s5_helpers_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r2s5_helpers.ncu-rep
python scripts/ncu_select.py gpurun_out/r2s5_helpers.ncu-rep > gpurun_out/r2s5_ncu_helpers_selected.csv
cut -c1-200 gpurun_out/r2s5_ncu_helpers_selected.csv | head -30
