"""experiment: does de-phasing help?  K independent training steps of 256/K windows each on K streams (K CUDA graphs
replayed concurrently) against one step of 256 windows.  Aggregate windows/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecg_denoise_b200 import synth
from ecg_denoise_b200.engine import FusedTrainer
from ecg_denoise_b200.model import transformer

def run(K, total=256, iters=40):
    B = total // K
    trs, xs, ts, streams = [], [], [], []
    for k in range(K):
        torch.manual_seed(2023 + k)
        m = transformer.ralenet(high_level_enhence=True).cuda()
        for rw in (m.rwattn1, m.rwattn2, m.rwattn3, m.rwattn4):
            rw.parameters_normalize()
        noisy, clean = synth.make_batch(B, 2, 256, seed=5 + k)
        xs.append(torch.from_numpy(noisy).cuda()); ts.append(torch.from_numpy(clean).cuda())
        trs.append(FusedTrainer(m, lr=1e-3, use_graph=True))
        streams.append(torch.cuda.Stream())
    for k in range(K):
        with torch.cuda.stream(streams[k]):
            for _ in range(3):
                trs[k].step(xs[k], ts[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                trs[k].step(xs[k], ts[k])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"K={K} x B={B}: {1e3 * dt / iters:.3f} ms per round of {total} windows -> {total * iters / dt:,.0f} windows/s", flush=True)

for total in (256, 512):
    for K in (1, 2, 4):
        run(K, total)
