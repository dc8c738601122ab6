#!/bin/bash
# decode bench under a list of "KNOB=value" settings (one run each); prints tokens/s and ms/step
for kv in "$@"; do
  env $kv timeout 600 python bench.py --steps ${STEPS:-32} --warmup 4 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('%-28s %8.2f tok/s  %7.3f ms/step  e2e %8.2f  launches/step %s' % ('$kv', r['value'], r['ms_per_step'], r['e2e']['value'], r.get('launches_per_step')))"
done
