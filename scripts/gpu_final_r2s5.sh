#!/bin/bash
# End-of-round evidence (session 5 of round 2): full GPU parity suite, smoke, default bench, kernel list + launch list of one eager step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2s5
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; echo "exit $?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "exit $?" >> gpurun_out/${T}_bench_default.err
tail -2 gpurun_out/${T}_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2s5_bench_default.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'parity_max_rel', 'gpu_launches')}, d.get('e2e', {}).get('value'), d['roofline'].get('frac'), d['roofline'].get('ms_per_launch'), d.get('clocks'), d.get('cpu_baseline'))
PY
STEP="python scripts/prof_step.py 32 hybrid16s"
timeout 300 $STEP > gpurun_out/${T}_step_kernels.txt 2>&1 &&
DF_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches.csv $STEP > gpurun_out/${T}_ncu_step.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${T}_launches.csv
python scripts/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_list_step.json
