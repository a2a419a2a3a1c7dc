#!/bin/bash
# same-box A/B of scripts/gemm_ab.py (hybrid16s) under an environment switch: gpu_env_ab.sh VAR A B [repeats]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
var=$1; a=$2; b=$3; n=${4:-2}
out=gpurun_out/env_ab.jsonl
: > $out
for i in $(seq $n); do
  for v in $a $b; do
    env $var=$v timeout 300 python scripts/gemm_ab.py hybrid16s 2>gpurun_out/env_ab.err | python -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin if l.startswith('{')]
print(json.dumps({'$var': '$v', 'ms': {r['case'][:22]: r['hybrid16s']['ms'] for r in rows}, 'err': {r['case'][:22]: r['hybrid16s'].get('err_vs_f64') for r in rows}}))" >> $out
  done
done
cat $out
tail -3 gpurun_out/env_ab.err
