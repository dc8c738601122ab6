#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_gemv_m8b.jsonl; rm -f $OUT
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --exact 0 --tc 0 --types q4 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -1; }
run --ms 1,2,4,5,8 --shapes 10240x5120,5120x8192,51200x5120,5120x25600
run --ms 1,4,8 --types q2t --shapes 12288x4096,4096x12288,6144x4096,4096x4096
timeout 900 python tools/throughput_bench.py --workload qwen3-32b-q4 --batch 2,4,8 --ctx 512 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('batch %d: %.1f tok/s' % (d['batch'], d['tokens_per_s']))" | tee -a $OUT
timeout 900 python tools/throughput_bench.py --workload qwen3-8b-q2 --batch 4,8 --ctx 512 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8b-q2 batch %d: %.1f tok/s' % (d['batch'], d['tokens_per_s']))" | tee -a $OUT
