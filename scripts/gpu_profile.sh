#!/bin/bash
# ncu evidence for one round: per-launch device times of a bench step + full captures of the dominant kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
for k in kv_fwd kv_dw xattn_fwd xattn_bwd; do
  pat="gemm_tc_kernel"; skip=1
  case $k in xattn_fwd) pat="xattn_fwd";; xattn_bwd) pat="xattn_bwd";; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o gpurun_out/prof_$k \
      python scripts/prof_kernels.py $k > gpurun_out/prof_$k.log 2>&1
  echo "$k exit $?"
done
ls -la gpurun_out/*.ncu-rep
