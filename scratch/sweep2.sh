#!/bin/bash
# usage: scratch/sweep2.sh "lib|threads" ...
for item in "$@"; do
  L=${item%%|*}; T=${item##*|}
  echo "LIB=$L THREADS=$T"
  DEXB200_THREADS=$T DEXB200_LIB=$L python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ms', round(d['ms_per_step'],4), 'nodeops/s %.3e' % d['value'], 'frac', round(d['roofline']['frac'],4))"
done
