#!/bin/bash
# first GPU pass: parity tests, microbench, a small bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q --maxfail=80 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/gemv_bench.py --quick --ms 1 --types q4,q2t,q1,f8,bf16 --splitk 0 --out gpurun_out/gemv_quick.jsonl > gpurun_out/gemv_quick.log 2>&1
tail -20 gpurun_out/gemv_quick.log
timeout 600 python bench.py --workload qwen3-0.6b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_06b.log 2>&1
tail -3 gpurun_out/bench_06b.log
