#!/bin/bash
# decode throughput of the three model sizes (batch 1) + the four 32B GEMV shapes: a quick before / after for prologue changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_quickbench.jsonl; rm -f $OUT
for w in qwen3-32b-q4 qwen3-8b-q1 qwen3-0.6b-h84; do
  timeout 600 python bench.py --workload $w --steps 64 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w: %.1f tok/s e2e %.1f frac %.3f gemv avg %.2f us' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']))" | tee -a $OUT
done
timeout 600 python tools/gemv_bench.py --types q4 --ms 1 --exact 0 --tc 0 --shapes 10240x5120,5120x8192,51200x5120,5120x25600,1280x5120,12800x5120 --out $OUT 2>&1 | grep -v '^{' | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemv or norm" 2>&1 | tail -1
