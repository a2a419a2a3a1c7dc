mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider --timeout=300 -k "ragged or streaming" > gpurun_out/pytest_rag.log 2>&1; echo "exit $?" >> gpurun_out/pytest_rag.log; tail -4 gpurun_out/pytest_rag.log
timeout 200 python scripts/prof_kernels.py 2 > gpurun_out/plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_q_kernel|loss_forward_kernel|knn1_d3_kernel' \
    -c 8 -o gpurun_out/prof_kernels python scripts/prof_kernels.py 1 > gpurun_out/ncu_prof.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
