#!/bin/bash
# experiment: launch priority of the data-gradient chain (capture stream, RALENET_MAIN_PRIO) against the forked
# weight-gradient branch (RALENET_SIDE_PRIO) inside the step graph.
# usage (through gpurun): bash tools/exp_prio.sh TAG            256 windows (twice, interleaved) and 4096 windows
#                         bash tools/exp_prio.sh TAG sweep ["B ..."]   512 / 1024 / 2048 windows (or the given batch sizes), base against main-hi
tag=${1:-prio}
o=gpurun_out/${tag}_exp_prio${2:+_$2}${3:+2}.txt
: > $o
B="--no-cpu-baseline --no-extras --no-profile"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py $B $EXTRA 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$name', '$EXTRA', d['ms_per_step'], round(d['value']), d['final_loss'])" | tee -a $o
}
python -c "import torch; print(torch.cuda.Stream.priority_range())" | tee -a $o
if [ "$2" = sweep ]; then
  for b in ${3:-512 1024 2048}; do
    EXTRA="--batch $b --steps 20 --warmup 4"
    run base RALENET_MAIN_PRIO=0
    run main-hi RALENET_MAIN_PRIO=-3
  done
  exit 0
fi
for rep in 1 2; do
  EXTRA="--steps 40 --warmup 5"
  run base RALENET_MAIN_PRIO=0
  run main-hi RALENET_MAIN_PRIO=-3
  run side-hi RALENET_MAIN_PRIO=0 RALENET_SIDE_PRIO=-3
done
EXTRA="--batch 4096 --steps 10 --warmup 3"
run base RALENET_MAIN_PRIO=0
run main-hi RALENET_MAIN_PRIO=-3
run side-hi RALENET_MAIN_PRIO=0 RALENET_SIDE_PRIO=-3
