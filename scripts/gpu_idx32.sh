#!/bin/bash
# 32-bit index decode in maxpool / im2col_s2: exact-equality helper test + encoder and stride-2 tests, then the kernels' times in one eager step
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_encoder_gpu.py -m gpu -q -p no:cacheprovider -x -k "helper or stride2 or encoder_vs_oracle" 2>&1 | tail -2 | tee gpurun_out/idx32_tests.txt
timeout 40 python scripts/prof_step.py 32 hybrid16s 2>/dev/null | grep -i "maxpool\|im2col_s2\|total device" | tee gpurun_out/idx32_kernels.txt
