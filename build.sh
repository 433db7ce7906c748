#!/bin/bash
# Build the C-ABI shared library in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
     ${MBQC_NVCC_EXTRA} -o mentpy_b200/_mbqc_b200.so mentpy_b200/csrc/mbqc_b200.cu
echo "built mentpy_b200/_mbqc_b200.so"
