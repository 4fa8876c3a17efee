#!/bin/bash
# multi-GPU evidence: the data-parallel equivalence check (symmetric-memory exchange kernels: multimem and peer
# variants, NCCL path) and the bench line at N GPUs.  usage: gpurun --gpus N -- bash tools/gpu_final_multi.sh TAG N
tag=${1:-final}; n=${2:-2}
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
{ DP_GRAPH=1 timeout 300 $TR --master-port 29611 tests/dp_check.py
  RALENET_COMM_MULTIMEM=0 DP_GRAPH=1 timeout 300 $TR --master-port 29612 tests/dp_check.py
  RALENET_COMM=nccl DP_GRAPH=1 timeout 300 $TR --master-port 29613 tests/dp_check.py; } 2>&1 | grep "dp_check" | tee $o/${tag}_dp_check_${n}gpu.txt
timeout 600 $TR --master-port 29614 bench.py --gpus $n --steps 20 --warmup 5 2> $o/${tag}_bench_${n}gpu.err | grep "^{" > $o/${tag}_bench_${n}gpu.json
RALENET_COMM=nccl timeout 600 $TR --master-port 29615 bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-profile 2> /dev/null | grep "^{" > $o/${tag}_bench_${n}gpu_nccl.json
for f in bench_${n}gpu bench_${n}gpu_nccl; do cut -c1-200 $o/${tag}_$f.json; done
