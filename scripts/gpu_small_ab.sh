#!/bin/bash
# pyramid pool / pyramid sum / patch gather, later generations (DF_ENC_V1=1 = the first): encoder parity tests, per-kernel times of one
# eager step per variant, pose bench A/B on the same box.  Output: gpurun_out/small_*
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
for cfg in "DF_ENC_V1=1" "DF_ENC_SUM=1 DF_ENC_POOL_SPLIT=2" "DF_ENC_SUM=2 DF_ENC_POOL_SPLIT=1"; do
  env $cfg timeout 300 python scripts/prof_step.py 32 hybrid16s 2>/dev/null | grep -i "pyramid\|gather_up\|total device" | sed "s/^/[$cfg] /"
done | tee gpurun_out/small_kernels.txt
for v in 1 0 1 0; do
  DF_ENC_V1=$v timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 3 2>> gpurun_out/small_ab.err | tail -1 > gpurun_out/small_line.json
  python - "$v" <<'P' | tee -a gpurun_out/small_ab.jsonl
import json, sys
d = json.loads(open('gpurun_out/small_line.json').read())
print(json.dumps({"DF_ENC_V1": int(sys.argv[1]), "poses_per_s": round(d["value"]), "e2e": round(d["e2e"]["value"]), "ms_per_step": round(d["ms_per_step"], 4),
                  "parity_max_rel": d.get("parity_max_rel"), "sm_mhz": d["clocks"]["sm_mhz"]}))
P
done
