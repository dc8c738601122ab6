#!/bin/bash
# tuning experiments: the same microbench against koifish_b200/variants/*.so builds (KF_TC_EXP bit flags)
for v in "" $(ls koifish_b200/variants/*.so 2>/dev/null); do
  echo "== ${v:-default}"
  KF_LIB_PATH=$v timeout 300 python tools/gemv_bench.py --shapes ${SHAPE:-51200x5120} --ms ${MS:-16} --types ${TYPES:-q4} --tc 1 --iters 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print('%-4s M=%-5d %8.1f us %7.0f GB/s %.3f %7.1f TF' % (r['type'], r['M'], r['us'], r['GBps'], r['frac_measured'], r['tflops']))"
done
