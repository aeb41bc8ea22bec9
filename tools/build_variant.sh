#!/bin/bash
# usage: tools/build_variant.sh NAME -DHOP_BATCH=4 ...   -> build_variants/libdfl_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude "$@" \
  -Wno-deprecated-gpu-targets -c deflate_compression_b200/csrc/dfl_kernels.cu -o build_variants/$name.o
nvcc -shared -o build_variants/libdfl_$name.so build_variants/$name.o deflate_compression_b200/csrc/build/deflate_compress.o \
  deflate_compression_b200/csrc/build/deflate_decompress.o -lcudart -Wno-deprecated-gpu-targets
