#!/bin/bash
# Build libspnerf_sm100a.so in-tree (nvcc cross-compiles sm_100a without a GPU).
set -e
cd "$(dirname "$0")/.."
SRC=sp-nerf_b200/csrc
OUT=sp-nerf_b200/lib
mkdir -p $OUT build
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA}"
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find $SRC include -name '*.h' -newer $o -o -name '*.cuh' -newer $o)" ]; then
    nvcc $FLAGS -c $f -o $o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT/libspnerf_sm100a.so build/*.o -gencode arch=compute_100a,code=sm_100a
echo built $OUT/libspnerf_sm100a.so
