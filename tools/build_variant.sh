#!/bin/bash
# Build an experiment variant of the library: tools/build_variant.sh <tag> <extra nvcc flags...>
# -> sp-nerf_b200/lib/variants/libspnerf_<tag>.so (selected at run time with SPNERF_LIB=<path>; experiments only)
set -e
cd "$(dirname "$0")/.."
TAG=$1; shift
SRC=sp-nerf_b200/csrc
OUT=sp-nerf_b200/lib/variants
mkdir -p $OUT build/$TAG
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DSPNERF_EXPERIMENTS $*"
pids=()
for f in $SRC/*.cu; do
  nvcc $FLAGS -c $f -o build/$TAG/$(basename ${f%.cu}).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT/libspnerf_$TAG.so build/$TAG/*.o -gencode arch=compute_100a,code=sm_100a
echo built $OUT/libspnerf_$TAG.so
