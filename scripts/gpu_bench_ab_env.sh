#!/bin/bash
# same-box A/B of the default bench under an environment switch: gpu_bench_ab_env.sh VAR A B [repeats]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
var=$1; a=$2; b=$3; n=${4:-2}
out=gpurun_out/bench_ab_env.jsonl
: > $out
for i in $(seq $n); do
  for v in $a $b; do
    env $var=$v timeout 600 python bench.py --no-cpu-baseline --no-extras --no-parity --steps 20 2>gpurun_out/bench_ab_env.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'$var': '$v', 'poses_s': round(d['value'], 1), 'ms_per_step': round(d['ms_per_step'], 4), 'e2e': round(d['e2e']['value'], 1), 'tw_ms_sum': d['roofline']['time_weighted']['ms_sum'], 'tw_frac': d['roofline']['time_weighted']['frac'], 'sm_mhz': d['clocks']['sm_mhz']}))" >> $out
  done
done
cat $out
