"""Per-kernel roofline table from a `bench.py --dump-kernels` file:
   python tools/kernel_roofline.py profiles/r2_v26_kernel_events.json 256 [profiles/r2_v26_kernel_events_B4096.json 4096 ...]
algorithmic FLOPs / bytes per launch are bench.py::algorithmic_work; peaks from MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_work, load_peaks

peaks = load_peaks()
ptf, phbm = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("hbm_gbs", 6500.0)
args = sys.argv[1:]
print(f"peaks: {ptf} TFLOP/s bf16 sustained, {phbm} GB/s HBM copy (MEASURED_PEAKS.json)\n")
for path, B in zip(args[0::2], args[1::2]):
    B = int(B)
    d = json.load(open(path))
    print(f"### {os.path.basename(path)} -- {B} windows per launch, un-graphed sum {d['per_step_ms_sum']:.3f} ms/step\n")
    print("| kernel | launches/step | us/launch | us/step | algorithmic GFLOP | algorithmic MB | TFLOP/s | % of bf16 peak | GB/s | % of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for k in d["kernels"]:
        f, nb = algorithmic_work(k["label"], B)
        t = k["avg_ms"] * 1e-3
        if f and nb:
            tf, gb = f / t / 1e12, nb / t / 1e9
            print(f"| {k['label']} | {k['launches_per_step']} | {k['avg_ms']*1e3:.1f} | {k['ms_per_step']*1e3:.1f} | {f/1e9:.2f} | {nb/1e6:.1f} | "
                  f"{tf:.1f} | {100*tf/ptf:.2f} | {gb:.0f} | {100*gb/phbm:.1f} |")
        else:
            print(f"| {k['label']} | {k['launches_per_step']} | {k['avg_ms']*1e3:.1f} | {k['ms_per_step']*1e3:.1f} | | | | | | |")
    print()
