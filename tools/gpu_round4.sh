#!/bin/bash
# GPU pass 4: parity after the cp.async ring rewrite, compact sweep, bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -12
rm -f gpurun_out/gemv_sweep.jsonl
timeout 900 python tools/gemv_bench.py --shapes 51200x5120,5120x25600,10240x5120,5120x8192 --ms 1 --types q4 --splitk 0,1,2,3,4 --variants 1,2 \
    --out gpurun_out/gemv_sweep.jsonl > gpurun_out/gemv_sweep.log 2>&1
timeout 600 python tools/gemv_bench.py --quick --ms 1 --types q2t,q1,f8,bf16 --splitk 0 --variants 1,2 --out gpurun_out/gemv_sweep.jsonl > gpurun_out/gemv_sweep2.log 2>&1
python - <<'EOF'
import json
best = {}
for l in open("gpurun_out/gemv_sweep.jsonl"):
    r = json.loads(l)
    k = (r["type"], r["N"], r["K"])
    print("%-4s %6dx%-6d S=%d v=%d %7.1f us %7.0f GB/s %.3f" % (r["type"], r["N"], r["K"], r["splitk"], r["variant"], r["us"], r["GBps"], r["frac_measured"]))
EOF
timeout 900 python bench.py --workload qwen3-32b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_32b.log 2>&1
python - <<'EOF'
import json
try:
    r = json.loads(open("gpurun_out/bench_32b.log").read().strip().splitlines()[-1])
    print("32B: %.1f tok/s  %.3f ms/step  e2e %.1f  gemv frac %.3f  gemv share %.2f launches/step %.0f" % (
        r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"], r["roofline"]["gemv_share_of_step"], r["launches_per_step"]))
except Exception as e:
    print("bench parse failed", e)
    print(open("gpurun_out/bench_32b.log").read()[-1500:])
EOF
