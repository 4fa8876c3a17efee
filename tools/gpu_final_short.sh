#!/bin/bash
# the short evidence set (when the GPU budget does not allow gpu_final.sh): parity suite, the default bench line, smoke().
# usage (through gpurun): bash tools/gpu_final_short.sh TAG
tag=${1:-short}
o=gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $o/${tag}_pytest_gpu.txt
cp $o/parity_ops.json $o/${tag}_parity_ops.json 2>/dev/null; cp $o/parity_net.json $o/${tag}_parity_net.json 2>/dev/null
timeout 200 python bench.py --steps 20 --warmup 5 --dump-kernels $o/${tag}_kernel_events.json > $o/${tag}_bench.json 2> $o/${tag}_bench.err; head -c 600 $o/${tag}_bench.json; echo
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $o/${tag}_smoke.txt
