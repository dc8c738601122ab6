#!/bin/bash
# one ncu --set full capture of the tcgen05 GEMM kernel: SHAPE / M / TYPE from the environment
mkdir -p gpurun_out
NAME=${NAME:-tc_q4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_gemm_tc_kernel -s 4 -c 1 -f -o gpurun_out/$NAME \
  python tools/gemv_bench.py --shapes ${SHAPE:-51200x5120} --ms ${M:-16} --types ${TYPE:-q4} --tc 1 --iters 4 > gpurun_out/ncu_$NAME.log 2>&1
tail -3 gpurun_out/ncu_$NAME.log
ls -la gpurun_out/$NAME.ncu-rep
