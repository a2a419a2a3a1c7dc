#!/bin/bash
# 2-GPU run: NCCL data-parallel parity test + inference / training bench lines at N=2 (launched as the driver does).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider --timeout=500 -s > gpurun_out/pytest_dp.log 2>&1
echo "exit $?" >> gpurun_out/pytest_dp.log; tail -n 6 gpurun_out/pytest_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?" >> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json; tail -n 3 gpurun_out/bench_2gpu.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
echo "exit $?" >> gpurun_out/bench_ref_2gpu.err; tail -c 700 gpurun_out/bench_ref_2gpu.json
for ph in estimator refiner; do
  timeout 300 $TR bench.py --workload train --phase $ph --gpus 2 --steps 10 --warmup 3 > gpurun_out/train_${ph}_2gpu.json 2> gpurun_out/train_${ph}_2gpu.err
  echo "exit $?" >> gpurun_out/train_${ph}_2gpu.err; cut -c1-400 gpurun_out/train_${ph}_2gpu.json; tail -n 2 gpurun_out/train_${ph}_2gpu.err
done
