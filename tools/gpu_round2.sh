#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=80 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/gemv_quick.jsonl
timeout 900 python tools/gemv_bench.py --quick --ms 1 --types q4,q2t,q1,f8,bf16 --splitk 0 --out gpurun_out/gemv_quick.jsonl > gpurun_out/gemv_quick.log 2>&1
tail -16 gpurun_out/gemv_quick.log
timeout 600 python bench.py --workload qwen3-0.6b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_06b.log 2>&1
tail -2 gpurun_out/bench_06b.log
timeout 900 python bench.py --workload qwen3-32b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_32b.log 2>&1
tail -2 gpurun_out/bench_32b.log
