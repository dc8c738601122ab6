#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_gemv_m8.jsonl; rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --exact 0 --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -2; }
run --ms 2,4,8 --variants 0,2 --shapes 10240x5120,5120x8192,51200x5120,5120x25600
run --ms 8 --variants 0,2 --shapes 51200x5120 --splitk 1,2,3,4
run --ms 8 --variants 0,2 --shapes 5120x25600 --splitk 8,11,16,20
