#!/bin/bash
# usage: profiles/tools/build_variant.sh <tag> <H_S> <extra nvcc flags...>   -> build/libslode_<tag>.so (only that shape's TU is recompiled)
tag=$1; shape=$2; shift; shift
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Istructured_latent_odes_b200/csrc -Xptxas -v "$@" -c structured_latent_odes_b200/csrc/slode_fixed_${shape}.cu -o build/var_${tag}.o 2>&1 | python profiles/tools/ptxas_regs.py | grep -E "fixed_|rror" | head -20
objs=$(ls build/obj/*.o | grep -v slode_fixed_${shape}.o | grep -v _O1.o)
nvcc -shared -o build/libslode_${tag}.so $objs build/var_${tag}.o -gencode arch=compute_100a,code=sm_100a && echo built build/libslode_${tag}.so
