"""debug: newrale at B=1280 against the reference modules run eagerly on the same GPU (debug tool only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecg_denoise_b200 import ops, synth
from ecg_denoise_b200.model import ralenet_12leads as TW
from oracle import synth_weights as SW, ref_loader

def rel(a, b):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / (b.abs() + b.pow(2).mean().sqrt())).max())

# 1. Conv1dFn alone at B = 1280
for (ci, co, act) in ((12, 6, True), (6, 2, True), (2, 6, True), (6, 12, False)):
    torch.manual_seed(ci)
    x = torch.randn(1280, ci, 256, device="cuda", requires_grad=True)
    w = (torch.randn(co, ci, 13, device="cuda") * 0.2).requires_grad_(True)
    b = torch.randn(co, device="cuda", requires_grad=True)
    g = torch.randn(1280, co, 256, device="cuda") * 1e-6
    y = ops.Conv1dFn.apply(x, w, b, act, 0.01)
    y.backward(g)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    yr = torch.nn.functional.conv1d(x, w, b, padding=6)
    if act: yr = torch.nn.functional.leaky_relu(yr, 0.01)
    yr.backward(g)
    print(f"conv {ci}->{co}: y {rel(got[0], yr):.1e} dx {rel(got[1], x.grad):.1e} dw {rel(got[2], w.grad):.1e} db {rel(got[3], b.grad):.1e}", flush=True)

# 2. whole newrale vs reference eager
R = ref_loader.load_reference()
sd = SW.make_newrale_state_dict(2023)
noisy, clean = synth.make_batch(128, 12, 256, seed=12, kind="bw", snr_db=2.0)
gains = np.random.RandomState(2).uniform(0.6, 1.4, size=(10, 1, 1, 1)).astype(np.float32)
x = torch.from_numpy((noisy[None] * gains).reshape(1280, 12, 256)).cuda()
t = torch.from_numpy((clean[None] * gains).reshape(1280, 12, 256)).cuda()
m = TW.newrale(TW.ralenet(high_level_enhence=True)); m.load_state_dict(sd); m = m.cuda().train()
ref = R.ralenet_12leads.newrale(R.quiet(R.ralenet_12leads.ralenet, high_level_enhence=True)); ref.load_state_dict(sd); ref = ref.cuda().double().train()
for B in (64, 256, 1280):
    xs, ts = x[:B], t[:B]
    xa = xs.clone().requires_grad_(True); ya = m(xa); torch.nn.functional.mse_loss(ya, ts).backward()
    xb = xs.double().clone().requires_grad_(True); yb = ref(xb); torch.nn.functional.mse_loss(yb, ts.double()).backward()
    print(f"B={B}: out {rel(ya, yb):.1e} dx {rel(xa.grad, xb.grad):.1e} " + " ".join(f"{n}:{rel(p.grad, q.grad):.1e}" for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()) if p.requires_grad), flush=True)
    m.zero_grad(); ref.zero_grad()
