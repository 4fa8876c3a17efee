set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
export RALENET_ATTN_UMMA=1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attn" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
echo "attn tests rc=$rc"
if [ "$rc" = "0" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  echo "full suite (tile kernels on) rc=${PIPESTATUS[0]}"
  RALENET_ATTN_UMMA=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/k_on.json > gpurun_out/bench_on.json 2> gpurun_out/bench_on.err
  RALENET_ATTN_UMMA=1 timeout 300 python bench.py --workload infer --steps 3 --warmup 3 > gpurun_out/infer_on.json 2> gpurun_out/infer_on.err
fi
RALENET_ATTN_UMMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/k_off.json > gpurun_out/bench_off.json 2> gpurun_out/bench_off.err
RALENET_ATTN_UMMA=0 timeout 300 python bench.py --workload infer --steps 3 --warmup 3 > gpurun_out/infer_off.json 2> gpurun_out/infer_off.err
tail -c 600 gpurun_out/bench_on.json gpurun_out/bench_off.json gpurun_out/infer_on.json gpurun_out/infer_off.json
