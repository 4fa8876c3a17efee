#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite (memcheck on whole-network + new-kernel tests, racecheck /
# synccheck on the per-kernel tests of the round-2 kernels).  Slow (10-50x): keep the selection small.
# usage (through gpurun): bash tools/sanitize.sh TAG
tag=${1:-san}
o=gpurun_out
SEL_NET="test_train_forward_backward_vs_reference_golden and rw_le or test_finetune_trainer_vs_reference_golden and False-graph or test_fused_block_forward"
SEL_OPS="test_conv13 or test_comm_kernels or test_device_synth_batch_properties and emb or test_wgrad_register_tile_kernel_vs_oracle and 5-3 or test_standalone_helper or test_ffn_block and 4-1"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "$SEL_NET" > $o/${tag}_memcheck_net.txt 2>&1; echo "memcheck net rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "$SEL_OPS" > $o/${tag}_memcheck_ops.txt 2>&1; echo "memcheck ops rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "test_conv13 and mma or test_wgrad_register_tile_kernel_vs_oracle and 5-3 or test_device_synth_batch_properties and emb or test_attn_block and True or test_ffn_block and 1-1 or test_patch_merge and 1" > $o/${tag}_racecheck_ops.txt 2>&1; echo "racecheck ops rc=$?"
for f in memcheck_net memcheck_ops racecheck_ops; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $o/${tag}_$f.txt | tail -4; done
