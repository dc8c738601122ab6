#!/bin/bash
# the 8-GPU pass: parity at 8 ranks, bench.py at 8 / 4 ranks (fused and stand-alone exchange), and the launch list of one TP=8 token
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-v1}
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 "$@"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/tp_check.py > gpurun_out/r2_tp_check_n8_$TAG.txt 2>&1; echo "tp_check rc=$?"
grep "tp_check" gpurun_out/r2_tp_check_n8_$TAG.txt | tail -4
for cfg in "8 1" "4 1"; do
  set -- $cfg
  KF_TP_FUSED=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $1 --steps 64 --warmup 8 > gpurun_out/r2_bench_tp$1_fused$2_$TAG.log 2>&1; echo "bench N=$1 fused=$2 rc=$?"
  grep '^{' gpurun_out/r2_bench_tp$1_fused$2_$TAG.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('N=%d fused=$2: %.1f tok/s  %.3f ms/step  launches/step %.0f  gemv avg %.2f us' % (d['n_gpus'], d['value'], d['ms_per_step'], d['launches_per_step'], d['roofline']['avg_launch_us']))"
done
KF_PROFILE=1 timeout 900 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
    bash tools/ncu_rank0.sh gpurun_out/r2_launches_tp8_$TAG.csv python bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_tp8_$TAG.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2_launches_tp8_$TAG.csv') if not l.startswith('==')))
hdr = rows[0]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
sc = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi: continue
    n = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '').split('<')[0]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', '')) * sc.get(r[ui], 1.0)
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/r2_launches_tp8_$TAG.txt', 'w') as f:
    f.write('# rank 0 of a TP=8 Qwen3-32B 4-bit decode, 2 tokens under ncu (gpu__time_duration.sum, kernels serialised, cold caches; peers run free)\n')
    for n, (c, t) in agg.items():
        line = '%-34s n=%4d  total %9.1f us  avg %7.2f us  share %.3f' % (n, c, t, t / c, t / tot)
        print(line); f.write(line + '\n')
    f.write('total %.1f us for 2 tokens\n' % tot)
PY
