#!/bin/bash
# DRAM traffic and duration of every launch of the timed decode steps (one ncu pass, serialised, cold caches); writes the per-kernel
# summary that bench.py reports as roofline.traffic
mkdir -p gpurun_out
KF_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/traffic_decode.csv python bench.py --workload ${WORKLOAD:-qwen3-32b-q4} --steps 2 --warmup 3 --no-cpu-baseline --ctx ${CTX:-512} \
    > gpurun_out/ncu_traffic.log 2>&1
python - <<'PY'
import csv, collections, json
rows = list(csv.reader(l for l in open('gpurun_out/traffic_decode.csv') if not l.startswith('==')))
hdr = rows[0]
ki, mi, ui, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value'), hdr.index('ID')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
for r in rows[1:]:
    if len(r) <= vi: continue
    n = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '').split('<')[0]
    per[n][r[mi]] += float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    if (r[ii], n) not in seen:
        seen.add((r[ii], n)); cnt[n] += 1
out = {}
for n, m in per.items():
    out[n] = {"launches": cnt[n], "avg_us": m['gpu__time_duration.sum'] / cnt[n],
              "avg_dram_read_bytes": m['dram__bytes_read.sum'] / cnt[n], "avg_dram_write_bytes": m['dram__bytes_write.sum'] / cnt[n]}
    print("%-28s n=%4d avg %8.2f us  read %10.3f MB  write %8.3f MB" % (n, cnt[n], out[n]['avg_us'], out[n]['avg_dram_read_bytes'] / 1e6, out[n]['avg_dram_write_bytes'] / 1e6))
import os, sys
sys.path.insert(0, '.')
import bench
json.dump({"kernel_digest": bench.kernel_digest(), "gemv_exact": int(os.environ.get("KF_GEMV_EXACT", "0")), "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over 2 decode steps of bench.py (KF_PROFILE=1), per launch averages", "kernels": out},
          open('gpurun_out/traffic_decode.json', 'w'), indent=1)
PY
