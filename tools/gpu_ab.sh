#!/bin/bash
# same-box A/B of the working-tree library against build/variants/base.so (tools/build_variant.sh base at the old commit):
# parity suite on the new library, then the train step (B = 256 and 4096) with the live kernel table for both.
# usage (through gpurun): bash tools/gpu_ab.sh TAG [ncu]
tag=${1:-ab}
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--no-cpu-baseline --no-extras"
for v in base new base new; do
  if [ $v = base ]; then export RALENET_B200_LIB=build/variants/base.so; else unset RALENET_B200_LIB; fi
  timeout 300 python bench.py --steps 30 --warmup 5 $B --dump-kernels $o/k_${tag}_$v.json 2> $o/bench_${tag}_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', 'B256', d['ms_per_step'], d['value'])"
done
for v in base new; do
  if [ $v = base ]; then export RALENET_B200_LIB=build/variants/base.so; else unset RALENET_B200_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --batch 4096 $B --dump-kernels $o/k_${tag}_${v}_B4096.json 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v', 'B4096', d['ms_per_step'], d['value'])"
done
unset RALENET_B200_LIB
python tools/cmp_kernels.py $o/k_${tag}_base.json $o/k_${tag}_new.json | head -24
if [ "$2" = ncu ]; then bash tools/ncu_step.sh $tag > /dev/null 2>&1; tail -1 $o/${tag}_ncu.log; fi
