#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each) -> gpurun_out/*.ncu-rep
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 300 $NCU -k regex:pool_foldw_fwd -s 1 -o gpurun_out/prof_foldw_fwd python scripts/fold_bench.py 2 > gpurun_out/prof_foldw_fwd.log 2>&1; echo "foldw fwd $?"
timeout 300 $NCU -k regex:pool_foldw_bwd -s 1 -o gpurun_out/prof_foldw_bwd python scripts/fold_bench.py 2 > gpurun_out/prof_foldw_bwd.log 2>&1; echo "foldw bwd $?"
timeout 300 $NCU -k regex:gemm_tc -s 2 -o gpurun_out/prof_gemm_split3 python scripts/prof_gemm_small.py split3 > gpurun_out/prof_gemm_split3.log 2>&1; echo "gemm split3 $?"
timeout 300 $NCU -k regex:gemm_tc -s 1 -o gpurun_out/prof_kv_fwd python scripts/prof_kernels.py kv_fwd > gpurun_out/prof_kv_fwd.log 2>&1; echo "kv fwd $?"
ls -la gpurun_out/*.ncu-rep
