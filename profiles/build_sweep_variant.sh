#!/bin/bash
# usage: build_sweep_variant.sh name "-DFLAGS"   -> fea-in-jax_b200/lib/variants/libfeb200_<name>.so (csrc/fused_sweep.cu rebuilt with FLAGS)
set -e
cd "$(dirname "$0")/../fea-in-jax_b200"
mkdir -p lib/variants build/var
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default $2 -c csrc/fused_sweep.cu -o build/var/fused_sweep_$1.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lib/variants/libfeb200_$1.so build/maps.o build/pattern.o build/element.o build/assemble.o build/fused.o build/fused_mma.o build/var/fused_sweep_$1.o build/krylov.o build/dist.o build/multigrid.o build/tangent_mma.o -lcudart -ldl
