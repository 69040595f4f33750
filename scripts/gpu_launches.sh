#!/bin/bash
# per-launch device times of two bench steps (ncu, serialised, cold cache) -> gpurun_out/launches.csv
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
