#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs under ncu (one pass, durations only -- kernels are serialised on that rank; the peers run free
# and wait for it at every exchange), the other ranks run the command as is.  Usage: ... --no-python bash tools/ncu_rank0.sh OUT.csv python bench.py ...
OUT=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
    exec ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT" "$@"
else
    exec "$@"
fi
