#!/bin/bash
# ring depth of the 4-bit decode GEMV: 3 (default build) against 4 (libkoifish_b200_d4.so), launch by launch and inside the decode step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_gemv_depth.jsonl; rm -f $OUT
for lib in "" koifish_b200/libkoifish_b200_d4.so; do
  echo "## lib=${lib:-default}" | tee -a $OUT
  KF_LIB_PATH=$lib timeout 600 python tools/gemv_bench.py --types q4 --ms 1 --exact 0 --tc 0 --shapes 10240x5120,5120x8192,51200x5120,5120x25600 --out $OUT 2>&1 | grep -v '^{' | tail -2
  KF_LIB_PATH=$lib timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('lib=${lib:-default}: %.1f tok/s  frac %.3f  gemv avg %.2f us' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']))" | tee -a $OUT
done
