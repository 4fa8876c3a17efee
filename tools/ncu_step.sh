#!/bin/bash
# `ncu --set full` of the dominant kernels of one eager training step (source-level), exported as CSV.
# Run on the GPU box: gpurun -- bash tools/ncu_step.sh <tag>.  A whole-step full capture (~130 kernels x 39 passes with a
# 466 MB workspace to save/restore per pass) takes > 15 min -- do not do that: the regex picks one launch family each.
tag=${1:-step}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --graph off --no-cpu-baseline --no-profile --no-extras"
RX='regex:(attn_bwd_kernel<.int.16,|attn_bwd_kernel<.int.8,|attn_bwd_kernel<.int.128,|block_fwd_kernel<.int.16|block_fwd_kernel<.int.8|ffn_bwd_kernel<.int.16,|ffn_fwd_umma_kernel<.int.128|ffn_bwd_umma_kernel<.int.128|attn_fwd_umma_kernel<.int.128|attn_fwd_umma_kernel<.int.64|wgrad_umma_kernel<.int.128|wgrad_reg_kernel|patch_bwd_kernel<.int.32,)'
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "$RX" -s 70 -c 22 \
  -o gpurun_out/${tag}_top $BENCH > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_top.ncu-rep --page raw --csv > gpurun_out/${tag}_top_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
