#!/bin/bash
# Build libwgg_sm100.so in-tree (sm_100a only).  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
ROOT="$(dirname "$PKG")"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$PKG/libwgg_sm100.so"
SRCS=(api.cu gemm.cu gemm_tc.cu ew.cu lstm.cu lstm_tc.cu conv_tc.cu encoder.cu disc.cu loss.cu optim.cu eval.cu keyboard.cu)
cd "$HERE"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared \
  -I"$ROOT/include" "$@" "${SRCS[@]}" -o "$OUT" -lcudart
echo "built $OUT"
