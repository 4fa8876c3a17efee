#!/bin/bash
# build a variant of libralenet_b200.so with extra nvcc defines: tools/build_variant.sh NAME -DRL_FW_MAXC=0 ...
# -> build/variants/NAME.so (select it with RALENET_B200_LIB=build/variants/NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name; mkdir -p $out
pids=()
for cu in ecg_denoise_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Iecg_denoise_b200/csrc "$@" -c $cu -o $out/$(basename ${cu%.cu}).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/$name.so $out/*.o
echo build/variants/$name.so
