#!/bin/bash
# one GPU iteration: parity suite, phase traces of the attention kernels, train-step bench with the live kernel table
# usage (through gpurun): bash tools/gpu_iter.sh TAG
tag=${1:-iter}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
if [ -f build/variants/trace.so ]; then
  RALENET_B200_LIB=build/variants/trace.so timeout 300 python tools/trace_attn.py > gpurun_out/trace_attn_$tag.txt 2>&1
  for c in 128 64; do for sv in 1 0; do RALENET_B200_LIB=build/variants/trace.so timeout 120 python tools/trace_attn_umma.py $c 256 $sv; done; done > gpurun_out/trace_attn_umma_$tag.txt 2>&1
  for c in 128 64; do RALENET_B200_LIB=build/variants/trace.so timeout 120 python tools/trace_umma.py $c; done > gpurun_out/trace_umma_$tag.txt 2>&1
fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/k_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cut -c1-330 gpurun_out/bench_$tag.json
