#!/bin/bash
# Development aid: compile the float/3D tile kernels alone and print registers + SASS instruction counts
# usage: tools/sass_stats.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v "$@" -o build/f3.cubin tools/sass_f3.cu 2>&1 \
  | grep -E "Compiling|Used|spill" | paste - - - | sed -E 's/ptxas info    : //g; s/Compiling entry function//; s/for .sm_100a.//' | cut -c1-220
for k in enc3_f32_tile_kernelILb0 dec3_f32_tile_kernelILb1ELb0; do
  f=$(cuobjdump -elf build/f3.cubin | grep -o "_ZN3zfb[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)
  cuobjdump -sass -fun "$f" build/f3.cubin > build/${k}.sass
  echo "$k: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' build/${k}.sass) SASS instructions"
done
