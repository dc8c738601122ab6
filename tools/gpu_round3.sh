#!/bin/bash
# GPU pass 3: parity after the GEMV v2 restructure, split-K / tile-variant sweep, bench, ncu launch list + one full capture
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/gemv_sweep.jsonl
timeout 900 python tools/gemv_bench.py --shapes 51200x5120,5120x25600,10240x5120,5120x8192,4096x4096 --ms 1 --types q4 --splitk 0,1,2,4,8 --variants 1,2 \
    --out gpurun_out/gemv_sweep.jsonl > gpurun_out/gemv_sweep.log 2>&1
tail -50 gpurun_out/gemv_sweep.log
timeout 600 python tools/gemv_bench.py --quick --ms 1 --types q2t,q1,f8,bf16 --splitk 0 --variants 1,2 --out gpurun_out/gemv_sweep.jsonl > gpurun_out/gemv_sweep2.log 2>&1
tail -24 gpurun_out/gemv_sweep2.log
timeout 900 python bench.py --workload qwen3-32b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_32b.log 2>&1
tail -2 gpurun_out/bench_32b.log
# every launch of two decode steps with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_32b.csv \
    python bench.py --workload qwen3-32b-q4 --steps 2 --warmup 3 --no-cpu-baseline --ctx 128 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
