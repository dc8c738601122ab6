#!/bin/bash
# ncu --set full captures of the decode attention kernels inside the Qwen3-32B decode step (cluster kernel at ctx 512, kv-group kernel at ctx 4096)
mkdir -p gpurun_out
KF_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kf_attn_cluster_kernel -s 8 -c 1 -f -o gpurun_out/attn_cluster \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ctx 512 > gpurun_out/ncu_attn_cluster.log 2>&1
KF_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kf_attn_gqa_kernel -s 8 -c 1 -f -o gpurun_out/attn_gqa \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ctx 4096 > gpurun_out/ncu_attn_gqa.log 2>&1
ls -la gpurun_out/attn_cluster.ncu-rep gpurun_out/attn_gqa.ncu-rep
