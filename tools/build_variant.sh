#!/bin/bash
# Experiment build of the library: history.cu recompiled with extra -D switches, the other objects as built by `make lib`.
# usage: tools/build_variant.sh NAME [-DFOO=1 ...]  ->  xmimsim_b200/lib/exp/lib_NAME.so  (use with XMIMSIM_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p xmimsim_b200/lib/exp build/exp
nvcc -ccbin g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -Ixmimsim_b200/csrc \
     -Xcompiler -fPIC,-fopenmp,-O3 "$@" -c xmimsim_b200/csrc/history.cu -o build/exp/history_$name.o
objs=$(ls build/obj/*.o | grep -v history.cu.o)
nvcc -ccbin g++ -gencode arch=compute_100a,code=sm_100a -shared -o xmimsim_b200/lib/exp/lib_$name.so build/exp/history_$name.o $objs \
     -Xcompiler -fopenmp -lgomp -ldl -cudart static
echo xmimsim_b200/lib/exp/lib_$name.so
