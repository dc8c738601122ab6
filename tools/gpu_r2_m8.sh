#!/bin/bash
# 2..8-token decode GEMV after round 2: pre-staged activations where the plan needs two waves + cluster merge of the k-slices up to 8 tokens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_gemv_m8.jsonl; rm -f $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2_m8_pytest.txt
run() { echo "## $*" | tee -a $OUT; timeout 600 python tools/gemv_bench.py --exact 0 --tc 0 --out $OUT "$@" 2>&1 | grep -v '^{' | tail -2; }
S=10240x5120,5120x8192,51200x5120,5120x25600
run --types q4 --ms 1,2,3,4,8 --shapes $S
run --types q4 --ms 2,4,8 --shapes $S --set gemv_xg_min_m=0,gemv_cluster=0
timeout 900 python tools/throughput_bench.py --workload qwen3-32b-q4 --batch 1,2,4,8 --ctx 512 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('batch %d: %.1f tok/s' % (d['batch'], d['tokens_per_s']))" | tee -a $OUT
timeout 900 python tools/throughput_bench.py --workload qwen3-8b-q2 --batch 4,8 --ctx 512 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('8b-q2 batch %d: %.1f tok/s' % (d['batch'], d['tokens_per_s']))" | tee -a $OUT
