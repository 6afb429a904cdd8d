#!/bin/bash
# usage: scratch/build_grad_variant.sh name -DDEX_GRAD_U=2 ...   -> scratch/libs/libdex_<name>.so
cd /root/repo/dynamicexpressions.jl_b200/csrc
name=$1; shift
mkdir -p /root/repo/scratch/libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v "$@" -c dex_grad.cu -o /tmp/grad_$name.o 2> /tmp/grad_$name.log && \
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o /root/repo/scratch/libs/libdex_$name.so ../lib/obj/dex_api.o ../lib/obj/dex_eval.o /tmp/grad_$name.o ../lib/obj/dex_flatten.o
grep -A2 "grad_kernelIfLi5" /tmp/grad_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
