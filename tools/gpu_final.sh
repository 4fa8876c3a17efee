#!/bin/bash
# round-end evidence set on ONE B200: parity suite, smoke, the default bench line (all sub-records) with the live kernel
# table, the two reference arms, the B = 4096 kernel table, the ncu launch list of an eager step.
# usage (through gpurun): bash tools/gpu_final.sh TAG        (multi-GPU lines: tools/gpu_final_multi.sh)
tag=${1:-final}
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $o/${tag}_pytest_gpu.txt
cp $o/parity_ops.json $o/${tag}_parity_ops.json; cp $o/parity_net.json $o/${tag}_parity_net.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $o/${tag}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 --dump-kernels $o/${tag}_kernel_events.json > $o/${tag}_bench.json 2> $o/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_ref.err
timeout 600 python bench.py --impl reference-gpu --steps 20 --warmup 5 > $o/${tag}_bench_reference_gpu_arm.json 2> $o/${tag}_refgpu.err
timeout 300 python bench.py --steps 10 --warmup 3 --batch 4096 --no-extras --no-cpu-baseline --dump-kernels $o/${tag}_kernel_events_B4096.json > $o/${tag}_bench_B4096.json 2> /dev/null
timeout 300 python tools/exp_two_streams.py > $o/${tag}_exp_two_streams.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --graph off --no-cpu-baseline --no-profile --no-extras > $o/${tag}_launches.log 2>&1
for f in bench bench_reference_arm bench_reference_gpu_arm bench_B4096; do tail -1 $o/${tag}_$f.json | cut -c1-220; done
