#!/bin/bash
# per-launch device times of one prefill (ncu serialises and runs cold: compare SHARES, not absolutes)
mkdir -p gpurun_out
KF_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_prefill.csv python tools/throughput_bench.py --workload ${WORKLOAD:-qwen3-32b-q4} --batch "" --prefill ${T:-4096} --panel ${P:-2048} --layers ${LAYERS:-4} \
    > gpurun_out/ncu_launch_prefill.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_prefill.csv') if not l.startswith('==')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    n = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:48]
    tot[n] += float(r[vi].replace(',', '')) / 1000.0
    cnt[n] += 1
allt = sum(tot.values())
print("launches %d  total %.1f us (serialised, cold)" % (sum(cnt.values()), allt))
for n, t in tot.most_common(14):
    print("%-50s n=%4d total=%9.1f us avg=%8.2f us share=%.3f" % (n, cnt[n], t, t / cnt[n], t / allt))
PY
tail -2 gpurun_out/ncu_launch_prefill.log
