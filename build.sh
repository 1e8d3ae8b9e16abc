#!/bin/bash
# Build libhrp_b200.so (sm_100a only) in-tree and the oracle's C pieces (none yet: the oracle is numpy/PyTorch-CPU).
set -e
ROOT="$(cd "$(dirname "$0")" && pwd)"
PKG="$ROOT/holistic-robot-pose-estimation-study_b200"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$PKG/lib" "$ROOT/build"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
OBJS=""
for f in "$PKG"/csrc/*.cu; do
  o="$ROOT/build/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$PKG/csrc" "$ROOT/include" -name '*.h' -newer "$o")" ]; then
    echo "[nvcc] $(basename "$f")"
    $NVCC $FLAGS -c "$f" -o "$o" 2> "$o.log" || { cat "$o.log"; exit 1; }
  fi
  OBJS="$OBJS $o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/lib/libhrp_b200.so" $OBJS -lcuda
echo "built $PKG/lib/libhrp_b200.so"
