#!/bin/sh
# Builds libsab200.so in-tree for sm_100a.  -fmad=false is part of the numerical contract.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${SAB200_OUT:-$HERE/../libsab200.so}"     # SAB200_OUT: an alternative build for A/B timing (load it with SAB200_LIB)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false \
    -Xcompiler -fPIC -shared "$@" -o "$OUT.tmp.$$" "$HERE/sab200.cu" -lcudart
mv -f "$OUT.tmp.$$" "$OUT"     # atomic: a snapshot taken meanwhile never sees a half-written library
