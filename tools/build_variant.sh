#!/bin/bash
# Build a variant of libspecgpu.so with extra nvcc flags on ONE source (timing ablations / experiments):
#   bash tools/build_variant.sh <name> <source.cu> <flags...>   ->  build/variants/libspecgpu_<name>.so
# Use it with SPECGPU_LIB=build/variants/libspecgpu_<name>.so.  The other objects come from the normal build (build/*.o).
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/variants
obj=$root/build/variants/${src%.cu}_$name.o
nvcc -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC "$@" \
     -c $root/spectrogram_enhancement_b200/csrc/$src -o $obj
others=$(ls $root/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $root/build/variants/libspecgpu_$name.so $obj $others -gencode arch=compute_100a,code=sm_100a -lcuda
echo $root/build/variants/libspecgpu_$name.so
