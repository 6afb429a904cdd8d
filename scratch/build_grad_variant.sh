#!/bin/bash
# usage: scratch/build_grad_variant.sh name NT [-D...]   -> scratch/libs/libdex_<name>.so
cd /root/repo/dynamicexpressions.jl_b200/csrc
name=$1; nt=$2; shift; shift
mkdir -p /root/repo/scratch/libs /tmp/gradinc_$name
python3 gen_grad_ptx.py --nt $nt --out /tmp/gradinc_$name/dex_grad_f32.inc > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v -DDEX_GRAD_THREADS=$nt "-DDEX_GRAD_INC=\"/tmp/gradinc_$name/dex_grad_f32.inc\"" "$@" -c dex_grad.cu -o /tmp/grad_$name.o 2> /tmp/grad_$name.log && \
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o /root/repo/scratch/libs/libdex_$name.so ../lib/obj/dex_api.o ../lib/obj/dex_eval.o /tmp/grad_$name.o ../lib/obj/dex_flatten.o
grep -A2 "grad_kernelIfLi5ELi1ELb0" /tmp/grad_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
