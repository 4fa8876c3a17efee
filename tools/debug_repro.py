"""debug: is the training-mode forward bit-reproducible across fresh trainers?  does it read uninitialised workspace?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecg_denoise_b200 import synth, _lib
from ecg_denoise_b200.engine import FusedTrainer
from ecg_denoise_b200.model import transformer
from oracle import synth_weights as SW
sd = SW.make_state_dict("rw", 1, 17)
for B in (6, 8, 16, 37, 256):
    noisy, clean = synth.make_batch(B, 2, 256, seed=61)
    x, t = torch.from_numpy(noisy).cuda(), torch.from_numpy(clean).cuda()
    outs = []
    for poison in (None, 0x00, 0xFF, 0x7F):
        m = transformer.ralenet(high_level_enhence=True); m.load_state_dict(sd); m = m.cuda()
        tr = FusedTrainer(m, lr=1e-3)
        tr._prepare(x.shape, x.device)
        if poison is not None:
            tr._ws.fill_(poison)
        loss, _, _, out = tr.step(x, t)
        outs.append((out.clone(), m._plan.flat_grad.clone(), loss.item()))
    for mode in (0, 2):
        prev = _lib.set_attn_umma(mode)
        m = transformer.ralenet(high_level_enhence=True); m.load_state_dict(sd); m = m.cuda()
        tr = FusedTrainer(m, lr=1e-3); tr._prepare(x.shape, x.device); tr._ws.fill_(0xFF)
        loss, _, _, out = tr.step(x, t)
        outs.append((out.clone(), m._plan.flat_grad.clone(), loss.item()))
        _lib.set_attn_umma(prev)
    base = outs[0]
    print(f"B={B}:", " ".join(f"[out {'==' if torch.equal(o[0], base[0]) else f'{(o[0]-base[0]).abs().max().item():.1e}'} "
                               f"grad {(o[1]-base[1]).abs().max().item() / base[1].abs().max().item():.1e} "
                               f"nan {int(torch.isnan(o[0]).sum())}/{int(torch.isnan(o[1]).sum())}]" for o in outs[1:]), flush=True)
