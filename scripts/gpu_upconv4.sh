#!/bin/bash
# upconv_finish: one-quad (3, default) against two-quads-per-thread (4) kernel, same box, bit-identity by hash; the dedicated test under 4
mkdir -p gpurun_out
for v in 3 4 3 4; do DF_UPCONV_SMEM=$v timeout 120 python scripts/upconv_ab.py >> gpurun_out/upconv4_ab.jsonl 2>> gpurun_out/upconv4_ab.err; done
python - <<'P'
import json
rows = [json.loads(l) for l in open('gpurun_out/upconv4_ab.jsonl')]
for r in rows:
    print(r['smem'], {k: v['ms'] for k, v in r.items() if k != 'smem'})
a, b = rows[0], rows[1]
print('bit-identical:', all(a[k]['hash'] == b[k]['hash'] and a[k]['checksum'] == b[k]['checksum'] for k in a if k != 'smem'))
P
DF_UPCONV_SMEM=4 timeout 200 python -m pytest tests/test_encoder_gpu.py -m gpu -q -p no:cacheprovider -x -k "upconv or helper or encoder_vs_oracle" 2>&1 | tail -2
