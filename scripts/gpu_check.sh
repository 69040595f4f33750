#!/bin/bash
# One GPU-box visit: parity tests (one process per file so a trapped kernel cannot poison the rest), smoke, bench.
# Everything is logged under gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_gemm test_gpu_fold test_gpu_fullsize test_gpu_blocks test_gpu_scl test_gpu_model test_gpu_dropin test_gpu_multi; do
  echo "=== $f ===" | tee -a gpurun_out/pytest_summary.txt
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/pytest_summary.txt
  tail -n 25 gpurun_out/$f.log | tee -a gpurun_out/pytest_summary.txt
done
echo "=== smoke ===" | tee -a gpurun_out/pytest_summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a gpurun_out/pytest_summary.txt
tail -n 8 gpurun_out/smoke.log | tee -a gpurun_out/pytest_summary.txt
if [ "${1:-}" != "nobench" ]; then
  echo "=== bench ===" | tee -a gpurun_out/pytest_summary.txt
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?" | tee -a gpurun_out/pytest_summary.txt
  tail -n 3 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
fi
