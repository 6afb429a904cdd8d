#!/bin/bash
TAG=${1:-g1}
ncu --set full --clock-control none --import-source on -k regex:grad_kernel -s 1 -c 1 -f -o gpurun_out/prof_grad_${TAG} python benchmarks/configs.py --only C3 --reps 1 > gpurun_out/ncu_grad_${TAG}.log 2>&1
ls -la gpurun_out | tail -3
