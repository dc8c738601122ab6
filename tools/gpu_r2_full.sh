#!/bin/bash
# full GPU pass: every -m gpu test, bench.py (default arithmetic and bit-exact), the GEMV sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v8}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu_$TAG.txt 2>&1
tail -5 gpurun_out/r2_pytest_gpu_$TAG.txt
S="10240x5120,5120x8192,51200x5120,5120x25600"
OUT=gpurun_out/r2_gemv_sweep_$TAG.jsonl
rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --shapes $S --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -3; }
run --ms 1,2,4,8 --exact 0
run --ms 1 --exact 0 --shapes 10240x5120 --splitk 0,2,4,8
run --ms 1 --exact 0 --shapes 5120x8192 --splitk 0,4,8,16
run --ms 1 --exact 1
( time timeout 900 python bench.py --steps 64 --warmup 8 ) > gpurun_out/r2_bench_$TAG.log 2>&1
tail -1 gpurun_out/r2_bench_$TAG.log | cut -c1-300
KF_GEMV_EXACT=1 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_${TAG}_exact.log 2>&1
tail -1 gpurun_out/r2_bench_${TAG}_exact.log | cut -c1-300
