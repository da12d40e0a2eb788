#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>  -> variants/lib<name>.so (same ABI)
# Only assemble_tiled.cu is recompiled (with -DTFEM_FAST_BUILD: the headline instantiation alone, a few
# seconds); the other objects come from the in-tree build (python -m pytorch_fem_solver_b200.build).
set -e
name=$1; shift
out=variants/$name; mkdir -p $out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DTFEM_FAST_BUILD "$@" \
  -I include -I pytorch_fem_solver_b200/csrc -c pytorch_fem_solver_b200/csrc/assemble_tiled.cu -o $out/assemble_tiled.o
lib=pytorch_fem_solver_b200/lib
nvcc -shared -o variants/lib$name.so $out/assemble_tiled.o $lib/geometry.o $lib/forms.o $lib/scatter.o $lib/interp.o $lib/sparse.o -cudart static
echo variants/lib$name.so
