#!/bin/bash
# quick GPU pass: GEMV tests + accuracy of the fast mode + microbench old kernel (fast / exact) vs TMA kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v7}
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemv or linear" ) > gpurun_out/r2_pytest_gemv.txt 2>&1
tail -3 gpurun_out/r2_pytest_gemv.txt
KF_GEMV_TMA=0 timeout 600 python tools/fast_mode_check.py > gpurun_out/r2_fast_mode_check_$TAG.txt 2>&1
tail -3 gpurun_out/r2_fast_mode_check_$TAG.txt
S="10240x5120,5120x8192,51200x5120,5120x25600"
OUT=gpurun_out/r2_gemv_sweep_$TAG.jsonl
rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --shapes $S --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -3; }
run --ms 1,2,4,8 --set gemv_tma=0 --exact 0
run --ms 1 --set gemv_tma=0 --exact 1
run --ms 1 --set gemv_tma=0,gemv_cluster=0 --exact 0
run --ms 1 --set gemv_tma=0,pdl=0 --exact 0
KF_GEMV_TMA=0 KF_GEMV_EXACT=0 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_${TAG}_fast.log 2>&1
tail -1 gpurun_out/r2_bench_${TAG}_fast.log | cut -c1-300
