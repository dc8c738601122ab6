#!/bin/bash
# one ncu --set full capture of the flash prefill attention kernel inside a 4-layer Qwen3-32B prefill
mkdir -p gpurun_out
KF_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kf_attn_prefill_kernel -s 3 -c 1 -f -o gpurun_out/attn_prefill \
  python tools/throughput_bench.py --workload qwen3-32b-q4 --batch "" --prefill 4096 --panel 2048 --layers 4 > gpurun_out/ncu_attn_prefill.log 2>&1
tail -2 gpurun_out/ncu_attn_prefill.log
ls -la gpurun_out/attn_prefill.ncu-rep
