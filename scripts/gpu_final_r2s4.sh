#!/bin/bash
# End-of-round evidence (last session of round 2): full GPU parity suite, smoke, default bench, launch list of one eager step, ncu --set
# full captures of the hot launch shapes.  Every ncu command runs only after the same command exited 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2s4
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=900 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; echo "exit $?" >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "exit $?" >> gpurun_out/${T}_bench_default.err
tail -2 gpurun_out/${T}_bench_default.err
STEP="python scripts/prof_step.py 32 hybrid16s"
timeout 300 $STEP > gpurun_out/${T}_step_kernels.txt 2>&1 &&
DF_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches.csv $STEP > gpurun_out/${T}_ncu_step.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${T}_launches.csv
python scripts/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_list_step.json
bash scripts/gpu_prof_cases.sh $T tower1 l40 l40b l41b conv6 up1
for c in tower1 l40 l40b l41b conv6 up1; do
  python scripts/ncu_select.py gpurun_out/${T}_$c.ncu-rep $c > gpurun_out/${T}_sel_$c.csv 2>/dev/null
done
head -1 gpurun_out/${T}_sel_tower1.csv > gpurun_out/${T}_ncu_full_selected.csv; sed -n 2p gpurun_out/${T}_sel_tower1.csv >> gpurun_out/${T}_ncu_full_selected.csv
for c in tower1 l40 l40b l41b conv6 up1; do sed -n 3p gpurun_out/${T}_sel_$c.csv >> gpurun_out/${T}_ncu_full_selected.csv; done
cat gpurun_out/${T}_ncu_full_selected.csv | cut -c1-300
rm -f gpurun_out/${T}_l40.ncu-rep gpurun_out/${T}_conv6.ncu-rep gpurun_out/${T}_up1.ncu-rep gpurun_out/${T}_l41b.ncu-rep
