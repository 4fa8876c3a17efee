"""debug: dx of the frozen core at several batch sizes vs the trainable core (same kernels, NULL gradient slots)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecg_denoise_b200.model import ralenet_12leads as TW
from oracle import synth_weights as SW

sd = SW.make_newrale_state_dict(2023)
core_sd = {k[5:]: v for k, v in sd.items() if k.startswith("rale.")}
for B in (3, 256, 511, 512, 1280):
    torch.manual_seed(B)
    x = torch.randn(B, 2, 256, device="cuda")
    g = torch.randn(B, 2, 256, device="cuda")
    res = []
    for frozen in (False, True):
        m = TW.ralenet(high_level_enhence=True)
        m.load_state_dict(core_sd)
        m = m.cuda().train()
        if frozen:
            for p in m.parameters():
                p.requires_grad_(False)
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        y.backward(g)
        res.append((y.detach(), xi.grad))
    ey = (res[0][0] - res[1][0]).abs().max().item()
    ed = (res[0][1] - res[1][1]).abs().max().item() / res[0][1].abs().max().item()
    print(f"B={B}: out diff {ey:.2e}, dx rel diff frozen vs trainable {ed:.2e}", flush=True)
