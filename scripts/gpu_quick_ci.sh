#!/bin/bash
# one-process GPU parity run + the default bench line (short form of gpu_ci.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 -x > gpurun_out/q_pytest.log 2>&1; echo "exit $?" >> gpurun_out/q_pytest.log; tail -4 gpurun_out/q_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "exit $?" >> gpurun_out/q_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_max_rel')}, d.get('e2e', {}).get('value'), d['roofline'].get('frac'), d['roofline'].get('ms_per_launch'), d['roofline'].get('time_weighted'), d.get('clocks'))
PY
tail -2 gpurun_out/q_bench.err
