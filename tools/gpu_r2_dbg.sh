#!/bin/bash
# timing experiments with the -DKF_DEBUG_KNOBS build (results of the kernels are garbage): where does the TMA GEMV lose its time?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export KF_LIB_PATH=$PWD/koifish_b200/libkoifish_b200_dbg.so
S="10240x5120,51200x5120,5120x25600"
OUT=gpurun_out/r2_gemv_dbg_${1:-v1}.jsonl
rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 300 python tools/gemv_bench.py --types q4 --shapes $S --tc 0 --ms 1 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -3; }
run --set gemv_tma=1
run --set gemv_tma=1,debug_skip=16
run --set gemv_tma=1,debug_skip=32
run --set gemv_tma=1,debug_skip=48
run --set gemv_tma=1,debug_skip=32 --exact 0
