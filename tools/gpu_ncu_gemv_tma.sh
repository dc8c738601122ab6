#!/bin/bash
# one full ncu capture of the TMA GEMV (fast mode) on the gate/up shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v6}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_tma -s 4 -c 1 -f -o gpurun_out/r2_gemv_tma_${TAG}_fast16 \
  python tools/gemv_bench.py --shapes 51200x5120 --ms 1 --types q4 --tc 0 --iters 4 --exact 0 --set gemv_tma=1,gemv_tma_warps=16,gemv_tma_smem_kb=208 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_tma -s 4 -c 1 -f -o gpurun_out/r2_gemv_tma_${TAG}_fast8occ2 \
  python tools/gemv_bench.py --shapes 51200x5120 --ms 1 --types q4 --tc 0 --iters 4 --exact 0 --set gemv_tma=1,gemv_tma_occ=2 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
