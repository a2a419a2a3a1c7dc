#!/bin/bash
# 2-GPU run: NCCL data-parallel parity test + training / inference bench lines at N=1 and N=2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider --timeout=500 -s > gpurun_out/pytest_dp.log 2>&1
echo "exit $?" >> gpurun_out/pytest_dp.log; tail -n 8 gpurun_out/pytest_dp.log
for ph in estimator refiner; do
  timeout 300 python bench.py --workload train --phase $ph --steps 10 --warmup 3 > gpurun_out/train_${ph}_1gpu.json 2> gpurun_out/train_${ph}_1gpu.err
  echo "exit $?" >> gpurun_out/train_${ph}_1gpu.err; cat gpurun_out/train_${ph}_1gpu.json; tail -n 3 gpurun_out/train_${ph}_1gpu.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --workload train --phase $ph --gpus 2 --steps 10 --warmup 3 > gpurun_out/train_${ph}_2gpu.json 2> gpurun_out/train_${ph}_2gpu.err
  echo "exit $?" >> gpurun_out/train_${ph}_2gpu.err; cat gpurun_out/train_${ph}_2gpu.json; tail -n 3 gpurun_out/train_${ph}_2gpu.err
done
