#!/bin/bash
# usage: variant_perf.sh <lib.so> ... ; runs the standard perf lines per library variant
for lib in "$@"; do
  echo "== $lib"
  FDTDX_B200_LIB=$lib python scripts/quick_perf.py 2>&1 | grep -E "xchunk=0 rows=0"
done
