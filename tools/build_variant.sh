#!/bin/bash
# build libzfb_<tag>.so with extra -D flags:  tools/build_variant.sh tag -DZFB_WS_P_WARPS=16 ...
tag=$1; shift
cd "$(dirname "$0")/.."
nvcc -shared -Xcompiler -fPIC -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo "$@" \
  -o vizaly-foresight_b200/libzfb_$tag.so vizaly-foresight_b200/csrc/zfb_api.cu
