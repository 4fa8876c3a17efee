#!/usr/bin/env bash
# Build libralenet_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
SRC=ecg_denoise_b200/csrc
OUT=ecg_denoise_b200/libralenet_b200.so
mkdir -p build
objs=()
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename "${f%.cu}").o
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ $SRC/common.cuh -nt "$o" ] || [ include/ralenet_b200.h -nt "$o" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
         -Iinclude -I$SRC "$@" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${objs[@]}"
echo "built $OUT"
