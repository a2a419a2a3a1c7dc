#!/bin/bash
# balanced schedule of the long-K convolutions on / off, same box (scripts/split_probe.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/split_ab.jsonl
for v in 1 0 1 0; do
  DF_TC_SPLIT=$v timeout 240 python scripts/split_probe.py >> gpurun_out/split_ab.jsonl 2>gpurun_out/split_err.log || echo "{\"failed\": \"DF_TC_SPLIT=$v\", \"rc\": $?}" >> gpurun_out/split_ab.jsonl
done
cat gpurun_out/split_ab.jsonl | cut -c1-260
tail -5 gpurun_out/split_err.log
