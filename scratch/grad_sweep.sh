#!/bin/bash
# usage: scratch/grad_sweep.sh lib1 lib2 ...   ("" = default lib): C3 (and a grad parity test) per library
for L in "$@"; do
  echo "LIB=$L"
  DEXB200_LIB=$L python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "grad or diff or deriv" 2>&1 | tail -2
  DEXB200_LIB=$L python benchmarks/configs.py --only C3 --reps 10 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' C3 ms', round(d['ms'],4), 'frac', round(d['hbm_roofline_frac'],4))"
done
