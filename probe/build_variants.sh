#!/bin/bash
# Builds experiment variants of librtat_b200.so into build/variants/ (each: one source recompiled with extra -D flags,
# linked with the regular objects).  usage: probe/build_variants.sh <name> <source.cu> <flags...>
set -e
name=$1; src=$2; shift 2
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c rtatblas_b200/csrc/$src.cu -o build/variants/${src}_$name.o
objs=$(ls build/obj/*.o | grep -v "/$src.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librtat_b200_$name.so $objs build/variants/${src}_$name.o -Xlinker -rpath,/usr/local/cuda/lib64
echo built build/variants/librtat_b200_$name.so
