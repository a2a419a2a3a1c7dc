#!/bin/bash
# Round-2 (second session) evidence: truncation calibration of hybrid16s, launch list of one eager step, full captures of the hot
# launch shapes in hybrid16s.  Every ncu command runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
DF_TC_BIAS_COMP=0,0,0,0 timeout 200 python scripts/trunc_probe.py > gpurun_out/r2s2_trunc_raw.jsonl 2>&1
timeout 200 python scripts/trunc_probe.py > gpurun_out/r2s2_trunc_comp.jsonl 2>&1
STEP="python scripts/prof_step.py 32 hybrid16s"
timeout 300 $STEP > gpurun_out/r2s2_step_kernels.txt 2>&1 &&
DF_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2s2_launches.csv $STEP > gpurun_out/r2s2_ncu_step.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2s2_launches.csv
python scripts/summarize_launches.py gpurun_out/r2s2_launches.csv > gpurun_out/r2s2_launch_list_step.json
bash scripts/gpu_prof_cases.sh r2s2 tower1 l40 l41 conv6 up1 l1
