#!/bin/bash
# pose bench at 32 / 48 / 64 frames per step (same box): does a larger step amortise wave quantisation and launch gaps?
mkdir -p gpurun_out
for f in 32 64 48 32 64; do
  timeout 300 python bench.py --frames $f --no-cpu-baseline --no-extras --no-parity --steps 10 --warmup 3 2>> gpurun_out/frames_ab.err | tail -1 >> gpurun_out/frames_ab.jsonl
done
python - <<'P'
import json
for l in open('gpurun_out/frames_ab.jsonl'):
    d = json.loads(l)
    print(d['config']['frames_per_gpu_per_step'], round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 3), d['clocks'])
P
