#!/bin/bash
# Run on the GPU box (under gpurun): parity tests in two processes (exact-fp32 path first, tensor-core path
# second so that a trap in the tcgen05 bring-up cannot poison the rest), smoke, short benches.
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
TC='tensor_core or 3xtf32 or tf32 or hybrid'
echo "== pytest (fp32 / non tensor-core) =="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 -k "not ($TC)" > gpurun_out/pytest_fp32.log 2>&1
echo "exit $?" >> gpurun_out/pytest_fp32.log
tail -n 25 gpurun_out/pytest_fp32.log
echo "== pytest (tensor-core) =="
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s -k "$TC" > gpurun_out/pytest_tc.log 2>&1
echo "exit $?" >> gpurun_out/pytest_tc.log
tail -n 40 gpurun_out/pytest_tc.log
echo "== pytest (hybrid16s with the A planes forced into shared memory on every tile width) =="
DF_TC_A_SMEM=1 timeout 600 python -m pytest tests/test_head_gpu.py tests/test_encoder_gpu.py -m gpu -q -p no:cacheprovider --timeout=500 -k "hybrid16s or f16s or conv" > gpurun_out/pytest_asmem.log 2>&1
echo "exit $?" >> gpurun_out/pytest_asmem.log
tail -n 4 gpurun_out/pytest_asmem.log
echo "== smoke =="
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 5 gpurun_out/smoke.log
echo "== bench fp32 =="
timeout 900 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "exit $?" >> gpurun_out/bench_fp32.err
tail -c 3000 gpurun_out/bench_fp32.json; tail -n 5 gpurun_out/bench_fp32.err
echo "== bench 3xtf32 =="
timeout 900 python bench.py --precision 3xtf32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench_3xtf32.err; echo "exit $?" >> gpurun_out/bench_3xtf32.err
tail -c 1500 gpurun_out/bench_3xtf32.json; tail -n 5 gpurun_out/bench_3xtf32.err
echo "== bench hybrid =="
timeout 900 python bench.py --precision hybrid --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err; echo "exit $?" >> gpurun_out/bench_hybrid.err
tail -c 1500 gpurun_out/bench_hybrid.json; tail -n 5 gpurun_out/bench_hybrid.err
echo "== bench (default: hybrid16s) =="
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err
tail -c 4000 gpurun_out/bench_default.json; tail -n 5 gpurun_out/bench_default.err
