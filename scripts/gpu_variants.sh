#!/bin/bash
# same-box A/B of library variants built into scripts/_bin/lib<V>.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/variants.jsonl
: > $out
export DF_AB_ONLY="tower1,tower3,conv5,up_2,pf conv2,layer2.1,layer4.0,conv6"
for v in "$@"; do
  cp scripts/_bin/lib$v.so densefusion_b200/libdensefusion_b200.so
  echo "{\"variant\": \"$v\"}" >> $out
  timeout 200 python scripts/knockout_probe.py >> $out 2>gpurun_out/var_err.log || echo "{\"failed\": \"$v\"}" >> $out
  timeout 300 python scripts/gemm_ab.py hybrid16s 2>>gpurun_out/var_err.log | python -c "
import sys, json
print(json.dumps({json.loads(l)['case'][:18]: json.loads(l)['hybrid16s']['ms'] for l in sys.stdin}))" >> $out
done
cat $out
