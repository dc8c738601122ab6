#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_gemv_splitk.jsonl
rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --tc 0 --ms 1 --out $OUT --exact 0 "$@" 2>&1 | grep -v '^{' | tail -3; }
run --shapes 10240x5120 --set gemv_tma=0 --splitk 0,2,3,4,5,6,8,10
run --shapes 5120x8192 --set gemv_tma=0 --splitk 0,4,6,8,10,12,16
run --shapes 5120x25600 --set gemv_tma=0 --splitk 0,8,10,11,12,16,20,25
run --shapes 51200x5120 --set gemv_tma=0 --splitk 0,1,2,3,4
for L in d5 d8; do
  if [ -f koifish_b200/libkoifish_b200_$L.so ]; then
    export KF_LIB_PATH=$PWD/koifish_b200/libkoifish_b200_$L.so
    run --shapes 10240x5120,5120x8192,51200x5120,5120x25600 --set gemv_tma=0
    KF_GEMV_TMA=0 KF_GEMV_EXACT=0 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_$L.log 2>&1
    tail -1 gpurun_out/r2_bench_$L.log | cut -c1-200
    unset KF_LIB_PATH
  fi
done
