#!/bin/bash
# per-launch device times of two bench steps (ncu, serialised) -> gpurun_out/launches.csv (cold caches: ncu flushes
# between kernels) and gpurun_out/launches_warm.csv (--cache-control none: L2 state carried over like in the real step)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense --eager > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/launches_warm.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense --eager > gpurun_out/bench_under_ncu_warm.log 2>&1
echo "warm launch list exit $?"
