#!/bin/bash
# A/B of the tensor-core GEMM variants: each variant's parity tests and timing probe in its own process, so that a
# trap in a bring-up kernel cannot poison the others.  Output: gpurun_out/gemm_ab_*.log
mkdir -p gpurun_out
for v in ${PARITY_VARIANTS:-5 6 7}; do
  echo "== parity variant $v =="
  DF_TEST_VARIANTS=$v timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q -p no:cacheprovider --timeout=200 -x -s \
      -k "tensor_core_variants or identity_layout or epilogues" > gpurun_out/gemm_ab_parity_v$v.log 2>&1
  echo "exit $?" >> gpurun_out/gemm_ab_parity_v$v.log
  tail -n 6 gpurun_out/gemm_ab_parity_v$v.log
done
for v in ${PROBE_VARIANTS:-5 6 7}; do
  echo "== probe variant $v =="
  timeout 200 python scripts/tc_probe.py $v > gpurun_out/gemm_ab_probe_v$v.log 2>&1
  echo "exit $?" >> gpurun_out/gemm_ab_probe_v$v.log
  cat gpurun_out/gemm_ab_probe_v$v.log
done
