"""Where does the end-to-end loop lose time against the device-resident loop?  Times variants of the loop (200 steps
each, wall clock around a final synchronize) at 256 windows: python tools/exp_e2e_gap.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ecg_denoise_b200.engine import FusedTrainer
from ecg_denoise_b200 import synth

dev = torch.device("cuda:0")
B, NB, N = 256, 4, 200
model, _ = bench.bench_state_dict()
model = model.to(dev)
noisy, clean = synth.make_batch(NB * B, 2, 256, seed=2023)
hx = torch.from_numpy(noisy).view(NB, B, 2, 256).pin_memory()
ht = torch.from_numpy(clean).view(NB, B, 2, 256).pin_memory()
dx, dt = hx.to(dev), ht.to(dev)
tr = FusedTrainer(model, lr=1e-3, use_graph=True)
for i in range(5):
    tr.step(dx[i % NB], dt[i % NB])
torch.cuda.synchronize()


def timed(name, body, pre=None):
    if pre:
        pre()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        body(i)
    torch.cuda.synchronize()
    print(f"{name:58s} {1e3 * (time.perf_counter() - t0) / N:8.4f} ms/step")


timed("replay only (static buffers untouched)", lambda i: tr._replay())
timed("step(device tensors): 2 D2D + replay", lambda i: tr.step(dx[i % NB], dt[i % NB]))
timed("step + .item() every step", lambda i: tr.step(dx[i % NB], dt[i % NB])[0].item())
timed("step_host (H2D on the main stream) + .item()", lambda i: tr.step_host(hx[i % NB], ht[i % NB]).item())
timed("step_host, no read", lambda i: tr.step_host(hx[i % NB], ht[i % NB]))


def b_prefetch(i):
    tr.step_host(hx[i % NB], ht[i % NB])
    tr.prefetch_host(hx[(i + 1) % NB], ht[(i + 1) % NB])


timed("prefetch_host + step_host, no read", b_prefetch, pre=lambda: tr.prefetch_host(hx[0], ht[0]))
pend = [None]


def b_full(i):
    l = tr.step_host(hx[i % NB], ht[i % NB])
    nxt = tr.read_async(l)
    tr.prefetch_host(hx[(i + 1) % NB], ht[(i + 1) % NB])
    if pend[0] is not None:
        pend[0]()
    pend[0] = nxt


timed("prefetch_host + step_host + read_async (lagged)", b_full, pre=lambda: tr.prefetch_host(hx[0], ht[0]))


def b_read_only(i):
    l = tr.step(dx[i % NB], dt[i % NB])[0]
    nxt = tr.read_async(l)
    if pend[0] is not None:
        pend[0]()
    pend[0] = nxt


pend[0] = None
timed("step(device) + read_async (lagged)", b_read_only)
