#!/bin/bash
# Every BASELINE.json configuration beyond the headline one, measured by ONE script into ONE file (gpurun_out/r02_configs.jsonl ->
# profiles/r02_configs.jsonl).  Each line carries "config" = the 1-based index of BASELINE.json's `configs` it belongs to.
#   config 2: Qwen3-0.6B hybrid 8/4-bit decode, batch 1 (bench.py contract line) and batch 1/32 at ctx 512
#   config 3: Qwen3-8B 2-bit / 1-bit decode batch 1..64 + 2K-token prefill
#   config 4: Qwen3-32B 4-bit batched decode + 4K-token prefill (the batch-1 line is the default bench.py run)
#   config 5: dequant-GEMV/GEMM sweep K,N in {4096,5120,25600} x bits {8,4,2,1} x M in {1,16,128,2048}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_configs.jsonl
rm -f $OUT
tag() { python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l); d['config'] = $1; print(json.dumps(d))" >> $OUT; }
timeout 600 python bench.py --workload qwen3-0.6b-h84 --steps 64 --warmup 8 2>/dev/null | tag 2
timeout 600 python tools/throughput_bench.py --workload qwen3-0.6b-h84 --batch 1,32 --ctx 512 2>/dev/null | tag 2
for w in qwen3-8b-q2 qwen3-8b-q1; do
    timeout 600 python bench.py --workload $w --steps 64 --warmup 8 --no-cpu-baseline 2>/dev/null | tag 3
    timeout 900 python tools/throughput_bench.py --workload $w --batch 1,2,4,8,16,32,64 --ctx 512 --prefill 2048 --panel 2048 2>/dev/null | tag 3
done
timeout 900 python tools/throughput_bench.py --workload qwen3-32b-q4 --batch 1,8,64 --ctx 512 --prefill 4096 --panel 2048 2>/dev/null | tag 4
timeout 1500 python tools/gemv_bench.py --types f8,q4,q2t,q1 --ms 1,16,128,2048 --exact 0 \
    --shapes 4096x4096,4096x5120,4096x25600,5120x4096,5120x5120,5120x25600,25600x4096,25600x5120,25600x25600 2>/dev/null | tag 5
wc -l $OUT
