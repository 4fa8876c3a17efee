"""per-kernel timing of the eval forward at a large batch (the inference regime: many waves of CTAs):
   python tools/profile_infer.py [B] -> table of kernels (CUDA events after every launch, ralenet_profile_*)"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecg_denoise_b200 import _lib
from ecg_denoise_b200.model import transformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
torch.manual_seed(2023)
model = transformer.ralenet(high_level_enhence=True)
for rw in (model.rwattn1, model.rwattn2, model.rwattn3, model.rwattn4):
    rw.parameters_normalize()
model = model.to(dev).eval()
x = torch.randn(B, 2, 256, device=dev)
lib = _lib.load()
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"eval forward B={B}: {ms:.3f} ms -> {B / ms * 1e3:.0f} windows/s")
    st = torch.cuda.current_stream().cuda_stream
    agg = {}
    for r in range(3):
        torch.cuda.synchronize()
        _lib.check(lib.ralenet_profile_begin(ctypes.c_void_p(st)))
        model(x)
        n_max = 4096
        labels = ctypes.create_string_buffer(n_max * 48)
        msb = (ctypes.c_float * n_max)()
        n = lib.ralenet_profile_end(labels, 48, msb, n_max)
        for i in range(n):
            lab = labels.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
            a = agg.setdefault(lab, [0, 0.0])
            a[0] += 1
            a[1] += msb[i]
tot = sum(a[1] for a in agg.values()) / 3
print(f"sum of kernels {tot:.3f} ms")
rows = sorted(((l, a[0] // 3, a[1] / a[0], a[1] / 3) for l, a in agg.items()), key=lambda r: -r[3])
for l, c, avg, s in rows:
    print(f"{s:8.4f} ms {100 * s / tot:5.1f}%  x{c:<3} {avg * 1e3:8.1f} us  {l}")
if len(sys.argv) > 2:
    json.dump({"B": B, "ms": ms, "kernels": [{"label": l, "launches": c, "avg_ms": a, "ms": s} for l, c, a, s in rows]},
              open(sys.argv[2], "w"), indent=1)
