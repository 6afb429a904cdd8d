#!/bin/bash
for c in 24 32 48 64 96 128 192; do
  echo "CHUNK_INSTR=$c"
  DEXB200_CHUNK_INSTR=$c python benchmarks/configs.py --only C2,C6 --reps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['config'], round(d['ms'],4), round(d['ms_median'],4), round(d['hbm_roofline_frac'],4))"
done
