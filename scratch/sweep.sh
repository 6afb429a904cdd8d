#!/bin/bash
# usage: scratch/sweep.sh lib1 lib2 ...  ("" = default lib)
for L in "$@"; do
  echo "LIB=$L"
  DEXB200_LIB=$L python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ms', round(d['ms_per_step'],4), 'nodeops/s %.3e' % d['value'], 'frac', round(d['roofline']['frac'],4), 'e2e ms', round(d['e2e']['ms_per_step'],3))"
done
