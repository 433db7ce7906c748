#!/bin/bash
# Build the C-ABI shared library in-tree for sm_100a (nvcc cross-compiles without a GPU).
# Translation units are compiled in parallel into build/obj and linked into one .so.
#   ./build.sh            all sources        ./build.sh sv_lean_host   only that source, then link
set -e
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${MBQC_NVCC_EXTRA}"
OUT=${MBQC_BUILD_OUT:-mentpy_b200/_mbqc_b200.so}
OBJ=${MBQC_BUILD_OBJ:-build/obj}
mkdir -p "$OBJ"
SRCS="mbqc_b200 sv_lean_host sv_jit_host probes peer"
ONLY=${1:-$SRCS}
pids=""
for s in $ONLY; do
    nvcc $FLAGS -c -o "$OBJ/$s.o" "mentpy_b200/csrc/$s.cu" &
    pids="$pids $!"
done
for p in $pids; do wait $p; done
objs=""
for s in $SRCS; do objs="$objs $OBJ/$s.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $objs -ldl
echo "built $OUT"
