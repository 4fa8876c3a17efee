#!/bin/bash
# compute-sanitizer over the kernels that changed after the v42 sanitizer run (v43 - v49: stem / head weight gradients,
# the fc1-epilogue GELU of the forward feed-forward, the pairwise 8-byte GEMM epilogues that every block kernel uses):
# racecheck on the per-kernel tests of all block kernels, memcheck on whole-network training steps.
# usage (through gpurun): bash tools/sanitize_final.sh TAG
tag=${1:-sanf}
o=gpurun_out
SEL_RACE="test_attn_block and True or test_ffn_block or test_stem or test_head or test_patch_merge or test_patch_separate or test_tcgen05_block_kernels_vs_oracle_large_batch and 4-1"
SEL_NET="test_train_forward_backward_vs_reference_golden or test_fused_block_forward or test_finetune_trainer_vs_reference_golden and False"
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL_RACE" > $o/${tag}_racecheck_ops.txt 2>&1; echo "racecheck ops rc=$?"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "$SEL_NET" > $o/${tag}_memcheck_net.txt 2>&1; echo "memcheck net rc=$?"
for f in racecheck_ops memcheck_net; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $o/${tag}_$f.txt | tail -4; done
