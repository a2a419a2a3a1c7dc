#!/bin/bash
# same-box A/B of the default bench under two argument sets: gpu_bench_ab_flag.sh "<args A>" "<args B>" [repeats]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
a=$1; b=$2; n=${3:-2}
out=gpurun_out/bench_ab_flag.jsonl
: > $out
for i in $(seq $n); do
  for v in "$a" "$b"; do
    timeout 600 python bench.py --no-cpu-baseline --no-extras --no-parity --steps 20 $v 2>gpurun_out/bench_ab_flag.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'args': '$v', 'poses_s': round(d['value'], 1), 'ms_per_step': round(d['ms_per_step'], 4), 'e2e': round(d['e2e']['value'], 1), 'sm_mhz': d['clocks']['sm_mhz']}))" >> $out
  done
done
cat $out
