#!/bin/bash
# N-GPU run (default 8) of the pose bench and of config C4 (both phases), launched as the driver does.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s2_bench_${N}gpu.json 2> gpurun_out/r2s2_bench_${N}gpu.err
echo "exit $?" >> gpurun_out/r2s2_bench_${N}gpu.err; cut -c1-260 gpurun_out/r2s2_bench_${N}gpu.json; tail -n 2 gpurun_out/r2s2_bench_${N}gpu.err
timeout 500 $TR bench.py --workload train --phase both --gpus $N --steps 10 --warmup 3 > gpurun_out/r2s2_train_${N}gpu.jsonl 2> gpurun_out/r2s2_train_${N}gpu.err
echo "exit $?" >> gpurun_out/r2s2_train_${N}gpu.err; cut -c1-330 gpurun_out/r2s2_train_${N}gpu.jsonl; tail -n 2 gpurun_out/r2s2_train_${N}gpu.err
