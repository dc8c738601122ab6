#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_kernel -s 6 -c 1 -f -o gpurun_out/gemv_q4_v4 \
  python tools/gemv_bench.py --shapes 51200x5120 --ms 1 --types q4 --splitk 1 --variants 1 --iters 4 > gpurun_out/ncu_gemv_q4.log 2>&1
ls -la gpurun_out/*.ncu-rep
