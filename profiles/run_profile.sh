#!/bin/bash
# Runs on the GPU box (under gpurun).  Usage: bash profiles/run_profile.sh <tag> [eval|grad|c4]
#   eval (default): launch list of the bench command, one full ncu capture of the eval kernel, then
#                   the bench line itself (never under a profiler) and the reference arm
#   grad:           one full ncu capture of the gradient kernel on C3
# (two calls: gpurun brings back at most 64 MiB per call and a capture with sources is ~30 MB)
TAG=${1:-r02}
WHAT=${2:-eval}
mkdir -p gpurun_out
if [ "$WHAT" = "grad" ]; then
    ncu --set full --clock-control none --import-source on -k regex:grad_kernel -s 2 -c 1 \
        -f -o gpurun_out/prof_grad_${TAG} python benchmarks/profile_target.py C3 > gpurun_out/ncu_grad_${TAG}.log 2>&1
    exit 0
fi
if [ "$WHAT" = "c4" ]; then   # the wide-input kernel on one C4 shard (10 features, depth-12 trees, 2^17 samples)
    ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 \
        -f -o gpurun_out/prof_c4_${TAG} python benchmarks/profile_target.py C4 > gpurun_out/ncu_c4_${TAG}.log 2>&1
    exit 0
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_eval_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ls -la gpurun_out
