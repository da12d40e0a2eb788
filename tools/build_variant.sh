#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>  -> gpurun_out/variants/lib<name>.so (same ABI)
set -e
name=$1; shift
out=variants/$name; mkdir -p $out
for f in geometry forms scatter interp sparse assemble_tiled; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
    -I include -I pytorch_fem_solver_b200/csrc -c pytorch_fem_solver_b200/csrc/$f.cu -o $out/$f.o &
done
wait
nvcc -shared -o variants/lib$name.so $out/*.o -cudart static
echo variants/lib$name.so
