#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v1}
( timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q ) > gpurun_out/r2_pytest_lowbit_$TAG.txt 2>&1; tail -4 gpurun_out/r2_pytest_lowbit_$TAG.txt
OUT=gpurun_out/r2_gemv_lowbit_$TAG.jsonl; rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -3; }
run --types q4 --ms 1 --exact 0 --variants 0,2 --shapes 10240x5120,5120x8192,51200x5120,5120x25600
run --types q2t,q1 --ms 1 --exact 0 --variants 0,2 --shapes 25600x5120,12288x4096
for w in qwen3-32b-q4 qwen3-8b-q2 qwen3-8b-q1; do
  timeout 600 python bench.py --workload $w --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_${w}_$TAG.log 2>&1
  grep '^{' gpurun_out/r2_bench_${w}_$TAG.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$w: %.1f tok/s e2e %.1f  frac %.3f gemv avg %.2f us' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']))"
done
