#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the eval kernel.
# Usage: bash profiles/run_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_eval_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
