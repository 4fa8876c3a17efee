BENCH="python bench.py --steps 3 --warmup 3 --graph off --no-cpu-baseline --no-profile"
RX='regex:(attn_bwd_kernel<.int.16,|attn_bwd_kernel<.int.8,)'
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "$RX" -s 8 -c 2 -o gpurun_out/v9_attn_bwd $BENCH > gpurun_out/v9_ncu.log 2>&1
ncu -i gpurun_out/v9_attn_bwd.ncu-rep --page raw --csv > gpurun_out/v9_attn_bwd_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
