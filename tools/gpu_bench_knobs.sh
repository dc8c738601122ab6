#!/bin/bash
# parity, then the 32B bench under a few knob settings (one line each)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -6
run() {
  env "$@" timeout 600 python bench.py --workload qwen3-32b-q4 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_tmp.log 2>&1
  python - "$*" <<'PY'
import json, sys
try:
    r = json.loads(open("gpurun_out/bench_tmp.log").read().strip().splitlines()[-1])
    print("%-28s %.1f tok/s  %.3f ms/step  e2e %.1f  gemv frac %.3f share %.2f launches %.0f" % (sys.argv[1], r["value"], r["ms_per_step"], r["e2e"]["value"], r["roofline"]["frac"], r["roofline"]["gemv_share_of_step"], r["launches_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/bench_tmp.log").read()[-600:])
PY
}
run KF_NONE=1
run KF_ATTN_SPLIT=16
run KF_ATTN_SPLIT=8
run KF_ATTN_SPLIT=4
run KF_PDL=0
cp gpurun_out/bench_tmp.log gpurun_out/bench_32b_nopdl.log
