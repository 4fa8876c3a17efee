#!/bin/bash
# second `ncu --set full` capture set: the kernels tools/ncu_step.sh leaves out (C = 8 / 32 / 64 block kernels, patch
# layers, stem / head / loss / Adam).  usage (through gpurun): bash tools/ncu_step2.sh TAG
tag=${1:-step2}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --graph off --no-cpu-baseline --no-profile --no-extras"
RX='regex:(attn_bwd_kernel<.int.8,|attn_bwd_kernel<.int.32,|attn_bwd_kernel<.int.64,|block_fwd_kernel<.int.8|block_fwd_kernel<.int.32|ffn_bwd_kernel<.int.8,|ffn_bwd_kernel<.int.32,|ffn_bwd_umma_kernel<.int.64|ffn_fwd_umma_kernel<.int.64|patch_fwd_kernel|patch_bwd_kernel|stem_|head_|mse_kernel|adam_kernel|reduce16|wgrad_kernel)'
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "$RX" -s 60 -c 48 \
  -o gpurun_out/${tag}_top2 $BENCH > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_top2.ncu-rep --page raw --csv > gpurun_out/${tag}_top2_raw.csv 2>/dev/null
tail -1 gpurun_out/${tag}_ncu2.log
# the report itself is too large to travel back next to other files (64 MiB limit): keep the digest
python tools/ncu_digest.py gpurun_out/${tag}_top2.ncu-rep 26 > gpurun_out/${tag}_top2_digest.txt 2>&1
rm -f gpurun_out/${tag}_top2.ncu-rep
