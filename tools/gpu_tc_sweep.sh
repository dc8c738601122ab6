#!/bin/bash
# tensor-core GEMM vs skinny GEMV across token counts; prints a compact table
OUT=gpurun_out/tc_sweep.jsonl
rm -f $OUT
timeout 1200 python tools/gemv_bench.py --shapes ${SHAPES:-5120x5120,25600x5120,5120x25600,51200x5120} --ms ${MS:-1,4,16,64,128,512,2048} --types ${TYPES:-q4,q2t,q1,f8,bf16} --tc ${TC:-0,1} --iters ${ITERS:-20} --out $OUT > gpurun_out/tc_sweep.log 2>&1
python - <<PY
import json
for l in open("$OUT"):
    r = json.loads(l)
    print("%-4s %6dx%-6d M=%-5d tc=%d sk=%d %9.1f us %7.0f GB/s %.3f  %7.1f TFLOP/s" % (r["type"], r["N"], r["K"], r["M"], r["tc"], r["splitk"], r["us"], r["GBps"], r["frac_measured"], r["tflops"]))
PY
tail -3 gpurun_out/tc_sweep.log
