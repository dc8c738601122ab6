#!/bin/bash
# round-2 evidence on ONE B200: part A = every -m gpu test, smoke(), bench.py (both arms); part B = ncu launch list / DRAM traffic of the decode
# step and one --set full capture of the dominant kernel.  Outputs under gpurun_out/ (copied into profiles/r02_* by hand).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PART=${1:-A}
if [ "$PART" = "A" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu.txt
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; tail -2 gpurun_out/r02_smoke.txt
  timeout 900 python bench.py --impl reference --steps 64 --warmup 8 > gpurun_out/r02_bench_32b_reference_arm.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_32b_reference_arm.json
  timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/r02_bench_32b.json 2> gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench_32b.json
  KF_GEMV_EXACT=1 timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_32b_exact.json 2>> gpurun_out/r02_bench.err; cut -c1-120 gpurun_out/r02_bench_32b_exact.json
else
  bash tools/gpu_traffic.sh > gpurun_out/r02_traffic.log 2>&1; tail -8 gpurun_out/r02_traffic.log
  cp gpurun_out/traffic_decode.json gpurun_out/r02_traffic_decode.json
  bash tools/gpu_launchlist.sh > gpurun_out/r02_launches_decode.txt 2>&1; tail -10 gpurun_out/r02_launches_decode.txt
  # one --set full capture per shape; only the CSV pages travel back (two .ncu-rep files exceed gpurun's 64 MiB), the large shape keeps its report
  for shp in 51200x5120 10240x5120; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_gemv_kernel -s 6 -c 1 -f -o /tmp/r02_gemv_q4_fast_$shp \
      python tools/gemv_bench.py --shapes $shp --ms 1 --types q4 --exact 0 --tc 0 --iters 4 > gpurun_out/r02_ncu_gemv_$shp.log 2>&1
    ncu -i /tmp/r02_gemv_q4_fast_$shp.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gemv_q4_fast_${shp}_raw.csv 2>/dev/null
    ncu -i /tmp/r02_gemv_q4_fast_$shp.ncu-rep --page details --csv > gpurun_out/r02_ncu_gemv_q4_fast_${shp}_details.csv 2>/dev/null
  done
  ncu -i /tmp/r02_gemv_q4_fast_51200x5120.ncu-rep --page source --csv > gpurun_out/r02_ncu_gemv_q4_fast_51200x5120_source.csv 2>/dev/null
  ls -la gpurun_out | tail -12
fi
