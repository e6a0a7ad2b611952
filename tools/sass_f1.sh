#!/bin/bash
# Development aid: compile the float/1D line kernels alone; registers, static SASS size and instruction mix
# usage: tools/sass_f1.sh [extra nvcc flags]
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v "$@" -o build/f1.cubin tools/sass_f1.cu 2>&1 \
  | grep -E "Compiling|Used" | paste - - | sed -E 's/ptxas info    : //g; s/Compiling entry function//; s/for .sm_100a.//' | cut -c1-200
for k in enc1_f32_kernel dec1_f32_kernel; do
  f=$(cuobjdump -elf build/f1.cubin | grep -o "_ZN3zfb[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)
  cuobjdump -sass -fun "$f" build/f1.cubin > build/${k}.sass
  echo "$k: $(grep -cE '^[[:space:]]+/\*[0-9a-f]{4}\*/' build/${k}.sass) SASS instructions; top opcodes:" \
    $(grep -E '^[[:space:]]+/\*[0-9a-f]{4}\*/' build/${k}.sass | sed -E 's/^[[:space:]]+\/\*[0-9a-f]+\*\/[[:space:]]+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | cut -d. -f1 | sort | uniq -c | sort -rn | head -12 | awk '{printf "%s:%s ", $2, $1}')
done
