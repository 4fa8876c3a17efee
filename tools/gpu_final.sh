#!/bin/bash
# round-end evidence set on one B200: parity suite, smoke, the three bench workloads, the reference arm, the live
# kernel table and the ncu launch list of an eager step.  usage (through gpurun): bash tools/gpu_final.sh TAG
tag=${1:-final}
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $o/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/${tag}_smoke.txt
timeout 600 python bench.py --dump-kernels $o/${tag}_kernel_events.json > $o/${tag}_bench.json 2> $o/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_ref.err
timeout 600 python bench.py --workload infer --steps 5 --warmup 3 > $o/${tag}_bench_infer.json 2> $o/${tag}_infer.err
timeout 600 python bench.py --workload train12 --steps 20 --warmup 3 > $o/${tag}_bench_train12.json 2> $o/${tag}_train12.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --graph off --no-cpu-baseline --no-profile > $o/${tag}_launches.log 2>&1
for f in bench bench_reference_arm bench_infer bench_train12; do cut -c1-260 $o/${tag}_$f.json; done
