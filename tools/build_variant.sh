#!/bin/bash
# Experimental build of the library with extra -D flags into build/exp/<name>.so (loaded through
# RPOOL_B200_LIB; the in-tree library is untouched).  usage: tools/build_variant.sh <name> <nvcc flags...>
name=$1; shift
mkdir -p build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared \
  -DRPOOL_BUILD_ID="\"exp-$name\"" "$@" -o build/exp/$name.so chainer-maskrcnn_b200/csrc/rpool_api.cu
