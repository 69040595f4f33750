#!/bin/bash
# A/B timing of tuning knobs on the bench step (no e2e / cpu legs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "MVF_SIDE_STREAM=1 MVF_GEMM_FILL=1" "MVF_SIDE_STREAM=1 MVF_GEMM_FILL=20" "MVF_SIDE_STREAM=1 MVF_GEMM_FILL=60" "MVF_SIDE_STREAM=0 MVF_GEMM_FILL=40" "MVF_SIDE_STREAM=0 MVF_GEMM_FILL=600"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', round(d['ms_per_step'],3), 'videos/s', round(d['value'],1), 'kv', round(d['roofline']['ms_per_launch'],3), 'dw', round(d['roofline']['weight_grad_gemm']['ms_per_launch'],3), 'loss', d['loss'])
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
