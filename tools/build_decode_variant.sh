#!/bin/bash
# Development aid: link a library whose decode_nms.o comes from another source file (A/B runs with
# tools/compare_decode_libs.py inside one gpurun call).   tools/build_decode_variant.sh <decode_nms source> <out.so> [nvcc -D flags]
set -e
SRC=$1; OUT=$2; shift 2
HERE=$(cd "$(dirname "$0")/.." && pwd)
B=$HERE/yolo_v1_b200/csrc/_build
mkdir -p "$(dirname "$OUT")"
cp "$SRC" $HERE/yolo_v1_b200/csrc/_variant_decode.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -fmad=false "$@" \
     -c $HERE/yolo_v1_b200/csrc/_variant_decode.cu -o /tmp/_variant_decode.o
rm -f $HERE/yolo_v1_b200/csrc/_variant_decode.cu
OBJS=$(ls $B/*.o | grep -v decode_nms.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJS /tmp/_variant_decode.o -cudart static -Xlinker --exclude-libs,ALL -Xlinker -Bsymbolic
echo built $OUT
