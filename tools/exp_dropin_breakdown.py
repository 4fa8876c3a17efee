"""Host-time breakdown of the stock training loop (denoise_train.py:51-57) on the drop-in module at 256 windows:
python tools/exp_dropin_breakdown.py   -> ms per phase (synchronised after every phase) and the un-synchronised step"""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ecg_denoise_b200 import synth
from ecg_denoise_b200.model import transformer

dev = torch.device("cuda:0")
_, sd = bench.bench_state_dict()
m = transformer.ralenet(high_level_enhence=True)
m.load_state_dict(sd)
m = m.to(dev).train()
noisy, clean = synth.make_batch(256, 2, 256, seed=2023)
x, t = torch.from_numpy(noisy).to(dev), torch.from_numpy(clean).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
acc = {}


def phase(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += host
    a[1] += tot
    return r


N = 30
for it in range(N + 5):
    if it == 5:
        acc.clear()
    phase("zero_grad", lambda: opt.zero_grad())
    out = phase("forward", lambda: m(x))
    loss = phase("mse_loss", lambda: F.mse_loss(out, t))
    phase("item", lambda: loss.item())
    phase("backward", lambda: loss.backward())
    phase("adam.step", lambda: opt.step())
for k, (h, tot) in acc.items():
    print(f"{k:12s} host {1e3 * h / N:7.3f} ms   host+device {1e3 * tot / N:7.3f} ms")
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(N):
    opt.zero_grad(); out = m(x); loss = F.mse_loss(out, t); loss.item(); loss.backward(); opt.step()
torch.cuda.synchronize()
print(f"un-synchronised loop: {1e3 * (time.perf_counter() - t0) / N:.3f} ms/step")
for kw in ({"fused": True}, {"foreach": True}, {"flat": True}):
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, **kw) if "flat" not in kw else __import__("ecg_denoise_b200.optim", fromlist=["Adam"]).Adam(m, lr=1e-3)
    for it in range(5):
        opt.zero_grad(); out = m(x); loss = F.mse_loss(out, t); loss.item(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(N):
        opt.zero_grad(); out = m(x); loss = F.mse_loss(out, t); loss.item(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    print(f"Adam({kw}): {1e3 * (time.perf_counter() - t0) / N:.3f} ms/step")

if os.environ.get("DROPIN_CPROFILE"):
    import cProfile, pstats
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    pr = cProfile.Profile()
    pr.enable()
    for it in range(N):
        opt.zero_grad(); out = m(x); loss = F.mse_loss(out, t); loss.item(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
