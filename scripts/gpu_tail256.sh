#!/bin/bash
# A/B of 256-wide tiles with a half-masked tail tile on tower-1 (N = 1920): correctness probe, then timing with the switch off / on
# in separate processes (the switch is read once).  Output: gpurun_out/tail256_*.jsonl
mkdir -p gpurun_out
for ts in 1 0; do
  DF_TC_TAIL256=2 DF_TC_TSTORE=$ts timeout 200 python scripts/tail256_probe.py >> gpurun_out/tail256_probe.jsonl 2> gpurun_out/tail256_probe.err
  echo "probe tstore=$ts exit $?"
done
cat gpurun_out/tail256_probe.jsonl
for rep in 1 2; do
  for v in 0 2; do
    echo "{\"DF_TC_TAIL256\": $v}" >> gpurun_out/tail256_ab.jsonl
    DF_TC_TAIL256=$v DF_AB_ONLY=tower1 timeout 200 python scripts/gemm_ab.py hybrid16s >> gpurun_out/tail256_ab.jsonl 2>> gpurun_out/tail256_ab.err
  done
done
cat gpurun_out/tail256_ab.jsonl
DF_TC_TAIL256=2 timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q -p no:cacheprovider -x -k "tile_widths or head_vs_oracle or crop_bias" 2>&1 | tail -4
