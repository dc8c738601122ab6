#!/bin/bash
# round 2 GPU pass: parity tests (incl. the reference-kernel pin), the decode GEMV microbench old vs TMA kernel, bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v4}
( time timeout 900 python -m pytest tests/test_gpu_refkernels.py -m gpu -q ) > gpurun_out/r2_pytest_refkernels.txt 2>&1
tail -3 gpurun_out/r2_pytest_refkernels.txt
( time timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu.txt
S="10240x5120,5120x8192,51200x5120,5120x25600"
OUT=gpurun_out/r2_gemv_sweep_$TAG.jsonl
rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --shapes $S --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -3; }
run --ms 1,2,4,8 --set gemv_tma=1
run --ms 1,8 --set gemv_tma=1,gemv_tma_occ=2
run --ms 1 --set gemv_tma=1,gemv_tma_depth=3
run --ms 1 --set gemv_tma=1,deq_fma=0
run --ms 1,8 --set gemv_tma=1 --exact 0
run --ms 1 --set gemv_tma=1,gemv_tma_occ=2 --exact 0
run --ms 1 --set gemv_tma=1,gemv_tma_depth=3 --exact 0
run --ms 1 --set gemv_tma=1,pdl=0 --exact 0
( time timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline ) > gpurun_out/r2_bench_$TAG.log 2>&1
tail -4 gpurun_out/r2_bench_$TAG.log | cut -c1-400
KF_GEMV_EXACT=0 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_${TAG}_factor.log 2>&1
tail -1 gpurun_out/r2_bench_${TAG}_factor.log | cut -c1-400
bash tools/gpu_r2_dbg.sh $TAG
