#!/bin/bash
# one full ncu capture of the TMA GEMV (exact and factored dequant) on the gate/up shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v3}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_tma -s 4 -c 1 -f -o gpurun_out/r2_gemv_tma_${TAG}_exact \
  python tools/gemv_bench.py --shapes 51200x5120 --ms 1 --types q4 --tc 0 --iters 4 --set gemv_tma=1 > gpurun_out/ncu_gemv_tma_exact.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_tma -s 4 -c 1 -f -o gpurun_out/r2_gemv_tma_${TAG}_factor \
  python tools/gemv_bench.py --shapes 51200x5120 --ms 1 --types q4 --tc 0 --iters 4 --exact 0 --set gemv_tma=1 > gpurun_out/ncu_gemv_tma_factor.log 2>&1
ls -la gpurun_out/*.ncu-rep
