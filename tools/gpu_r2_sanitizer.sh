#!/bin/bash
# compute-sanitizer over the kernels that are new or changed in round 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer.txt; : > $OUT
run() { echo "\$ compute-sanitizer $*" >> $OUT; timeout 1500 compute-sanitizer "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | tail -6 >> $OUT; }
run --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fast or sampler or nf4 or axb or splitk"
run --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "fast and (decode_logits or qwen3_32b or long_context or decode_loop)"
run --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "sampler or nf4_linear or (fast_matches and 1040)"
cat $OUT
