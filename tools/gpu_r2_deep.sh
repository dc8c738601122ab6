#!/bin/bash
# deep ring for short k-slices (knob gemv_deep_steps): launch by launch over the k-split, and inside the decode step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_gemv_deep.jsonl; rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --types q4 --ms 1 --exact 0 --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -2; }
for d in 0; do
  run --set gemv_deep_steps=$d --shapes 10240x5120 --splitk 0,4,8
  run --set gemv_deep_steps=$d --shapes 5120x8192 --splitk 0,8,16
  run --set gemv_deep_steps=$d --shapes 5120x25600 --splitk 0,11,22
  run --set gemv_deep_steps=$d --shapes 1280x5120,5120x1024,12800x5120,5120x3200
done
run --set gemv_deep_steps=10 --ms 2,8 --shapes 10240x5120,5120x8192
run --set gemv_deep_steps=0 --ms 2,8 --shapes 10240x5120,5120x8192
for d in 0; do
  KF_GEMV_DEEP_STEPS=$d timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('deep_steps=$d: %.1f tok/s  frac %.3f  gemv avg %.2f us' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']))" | tee -a $OUT
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemv or linear" 2>&1 | tail -2
