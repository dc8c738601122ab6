#!/bin/bash
# tensor-parallel pass on N GPUs of one box: parity (tools/tp_check.py) and bench.py with the fused and the stand-alone exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-v1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/tp_check.py > gpurun_out/r2_tp_check_n${N}_$TAG.txt 2>&1; echo "tp_check rc=$?"
grep "tp_check" gpurun_out/r2_tp_check_n${N}_$TAG.txt | tail -5
for f in 1 0; do
  KF_TP_FUSED=$f timeout 900 $TR bench.py --gpus $N --steps 64 --warmup 8 > gpurun_out/r2_bench_tp${N}_fused${f}_$TAG.log 2>&1; echo "bench fused=$f rc=$?"
  grep '^{' gpurun_out/r2_bench_tp${N}_fused${f}_$TAG.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('N=%d fused=$f: %.1f tok/s  %.3f ms/step  launches/step %.0f  gemv avg %.2f us' % (d['n_gpus'], d['value'], d['ms_per_step'], d['launches_per_step'], d['roofline']['avg_launch_us']))"
done
