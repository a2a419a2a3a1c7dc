#!/bin/bash
# 2-GPU check after the helper-kernel changes: NCCL data-parallel parity test, pose bench and estimator training line at N=2.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider --timeout=300 -s > gpurun_out/r2s5_pytest_dp_2gpu.log 2>&1
echo "exit $?" >> gpurun_out/r2s5_pytest_dp_2gpu.log; tail -n 4 gpurun_out/r2s5_pytest_dp_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s5_bench_2gpu.json 2> gpurun_out/r2s5_bench_2gpu.err
echo "exit $?" >> gpurun_out/r2s5_bench_2gpu.err; tail -1 gpurun_out/r2s5_bench_2gpu.json | cut -c1-300; tail -n 1 gpurun_out/r2s5_bench_2gpu.err
timeout 300 $TR bench.py --workload train --phase estimator --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s5_train_est_2gpu.json 2> gpurun_out/r2s5_train_est_2gpu.err
echo "exit $?" >> gpurun_out/r2s5_train_est_2gpu.err; tail -1 gpurun_out/r2s5_train_est_2gpu.json | cut -c1-300; tail -n 1 gpurun_out/r2s5_train_est_2gpu.err
