"""debug: core dx at several batch sizes vs the fp64 eager reference on the GPU; where is the error?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecg_denoise_b200 import synth
from ecg_denoise_b200.model import transformer as T
from oracle import synth_weights as SW, ref_loader

R = ref_loader.load_reference()
sd = SW.make_state_dict("rw", 1, 2023)
noisy, clean = synth.make_batch(320, 2, 256, seed=12, kind="bw", snr_db=2.0)
x = torch.from_numpy(np.tile(noisy, (4, 1, 1))).cuda(); t = torch.from_numpy(np.tile(clean, (4, 1, 1))).cuda()
x = x * torch.linspace(0.6, 1.4, 1280, device="cuda").view(-1, 1, 1)
for gs in (1.0, 1e-3):
  for B in (8, 64, 100, 128, 256, 512, 1280):
    m = T.ralenet(high_level_enhence=True); m.load_state_dict(sd); m = m.cuda().train()
    ref = R.quiet(R.transformer.ralenet, high_level_enhence=True); ref.load_state_dict(sd); ref = ref.cuda().double().train()
    xs, ts = x[:B], t[:B]
    xa = xs.clone().requires_grad_(True); ya = m(xa); (gs * torch.nn.functional.mse_loss(ya, ts)).backward()
    xb = xs.double().clone().requires_grad_(True); yb = ref(xb); (gs * torch.nn.functional.mse_loss(yb, ts.double())).backward()
    d = (xa.grad.double() - xb.grad).abs()
    rms = xb.grad.pow(2).mean().sqrt()
    perwin = d.flatten(1).max(1).values / rms
    worst = int(perwin.argmax())
    pos = int(d[worst].flatten().argmax())
    print(f"gs={gs} B={B}: dx max err/rms {float(perwin.max()):.2e} median-window {float(perwin.median()):.2e} worst window {worst} pos {pos} "
          f"(ref {float(xb.grad[worst].flatten()[pos]):.3e} got {float(xa.grad[worst].flatten()[pos]):.3e}); #windows>1e-3: {int((perwin > 1e-3).sum())}", flush=True)
