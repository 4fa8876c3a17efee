"""Generate tests/golden/*.npz by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden [--ref /root/reference] [--out tests/golden]

The reference ships no tests or golden vectors for this path, so the parity
pin is "outputs of the reference itself run here".  The reference modules are imported from
`--ref` (never copied), loaded with the seeded synthetic weights of `oracle/synth_weights.py`
(`load_state_dict(strict=True)` -- which also proves the restated state_dict layout is exact),
and run on seeded synthetic ECG windows, in float64 (sharp pin) and float32 (what users run).

`model/ralenet_12leads.py` does not import as shipped (it ends in a body-less
`if __name__ == "__main__":`); it is exec'd from source with `pass` appended in memory.

Large gradient tensors are stored as a strided subsample plus their sum and L2 norm to keep the
fixtures small.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecg_denoise_b200 import synth  # noqa: E402
from oracle import synth_weights  # noqa: E402

SUB = 32           # tensors with more elements than this are subsampled
STRIDE_PRIME = 37


def subsample(t: torch.Tensor) -> np.ndarray:
    f = t.detach().reshape(-1).to(torch.float64).numpy()
    if f.size <= SUB:
        return f.copy()
    idx = (np.arange(SUB) * STRIDE_PRIME * (f.size // SUB // STRIDE_PRIME + 1)) % f.size
    return f[idx].copy()


def pack(t: torch.Tensor) -> np.ndarray:
    """[sum, l2 norm, subsample...] of a tensor, float64."""
    t = t.detach().to(torch.float64)
    return np.concatenate([[t.sum().item(), t.norm().item()], subsample(t)])


def import_reference(ref_root: str):
    sys.path.insert(0, ref_root)
    with contextlib.redirect_stdout(io.StringIO()):
        from model import transformer, raletransformer  # type: ignore
    src = open(os.path.join(ref_root, "model", "ralenet_12leads.py")).read() + "\n    pass\n"
    mod = types.ModuleType("ralenet_12leads_patched")
    exec(compile(src, "ralenet_12leads.py", "exec"), mod.__dict__)
    sys.path.insert(0, os.path.join(ref_root))
    from local_utils import evaluate  # type: ignore
    return transformer, raletransformer, mod, evaluate


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):   # reference Mlp.__init__ prints a flag
        return fn(*a, **k)


def run_case(model, sd, x, target, out: dict, tag: str, evaluate, trainable=None):
    """fwd (eval, fp64 & fp32), fwd+bwd (train, fp64), running stats, 3 Adam steps (fp32)."""
    # layout proof
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), (tag, [k for k in ref_sd if k not in sd][:5],
                                                    [k for k in sd if k not in ref_sd][:5])
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape) and ref_sd[k].dtype == sd[k].dtype, k
    model.load_state_dict(sd, strict=True)

    m64 = model.double()
    x64, t64 = x.double(), target.double()
    m64.eval()
    with torch.no_grad():
        out[f"{tag}/eval_out64"] = m64(x64).numpy()
    m64.train()
    m64.zero_grad()
    xg = x64.clone().requires_grad_(True)
    y = m64(xg)
    loss = torch.nn.functional.mse_loss(y, t64)
    loss.backward()
    out[f"{tag}/train_out64"] = y.detach().numpy()
    out[f"{tag}/loss64"] = np.float64(loss.item())
    out[f"{tag}/dx64"] = xg.grad.numpy()
    names, packed = [], []
    for k, p in m64.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        packed.append(pack(p.grad))
    out[f"{tag}/grad_names"] = np.array(names)
    out[f"{tag}/grads"] = np.concatenate(packed)      # per name: [sum, l2norm, samples...]
    bn = "rale.conv1.2." if any(k.startswith("rale.") for k in sd) else "conv1.2."
    new_sd = m64.state_dict()
    out[f"{tag}/bn_running_mean"] = new_sd[bn + "running_mean"].numpy()
    out[f"{tag}/bn_running_var"] = new_sd[bn + "running_var"].numpy()
    out[f"{tag}/bn_nbt"] = np.int64(new_sd[bn + "num_batches_tracked"].item())

    # fp32: what the user runs
    m32 = model.float()
    m32.load_state_dict(sd, strict=True)
    m32.eval()
    with torch.no_grad():
        y32 = m32(x)
    out[f"{tag}/eval_out32"] = y32.numpy()
    out[f"{tag}/eval_snr32"] = evaluate.SNR(target, y32).numpy()
    out[f"{tag}/eval_rmse32"] = evaluate.RMSE(target, y32).numpy()

    # 3-step Adam trajectory in fp64 (denoise_train.py:24, 51-57)
    m64 = model.double()
    m64.load_state_dict(sd, strict=True)
    m64.train()
    opt = torch.optim.Adam([p for p in m64.parameters() if p.requires_grad], lr=0.001)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        l = torch.nn.functional.mse_loss(m64(x64), t64)
        losses.append(l.item())
        l.backward()
        opt.step()
    out[f"{tag}/adam_losses64"] = np.array(losses)
    fin = m64.state_dict()
    out[f"{tag}/adam_params"] = np.concatenate([pack(fin[k]) for k in names])
    model.float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    transformer, raletransformer, twelve, evaluate = import_reference(a.ref)
    os.makedirs(a.out, exist_ok=True)
    torch.manual_seed(2023)
    B = 4
    noisy, clean = synth.make_batch(B, 2, 256, seed=2023, kind="emb", snr_db=-4.0)
    x, tgt = torch.from_numpy(noisy), torch.from_numpy(clean)
    out = {"x": noisy, "target": clean}

    # case rw_le: main.py index 4  -- ralenet(high_level_enhence=True)
    run_case(quiet(transformer.ralenet, high_level_enhence=True),
             synth_weights.make_state_dict("rw", 1, 2023), x, tgt, out, "rw_le", evaluate)
    # case rw_mlp: main.py index 3 -- ralenet(low_level_enhence=False) == plain MLP + R-wave bias
    run_case(quiet(transformer.ralenet, low_level_enhence=False),
             synth_weights.make_state_dict("rw", 0, 2024), x, tgt, out, "rw_mlp", evaluate)
    # case nra: main.py index 2 -- raletransformer.ralenet()
    run_case(quiet(raletransformer.ralenet),
             synth_weights.make_state_dict("nra", 1, 2025), x, tgt, out, "nra", evaluate)
    # case nra512: the only reference model that accepts 2 x 512 windows
    n5, c5 = synth.make_batch(2, 2, 512, seed=77, kind="ma", snr_db=0.0)
    out["x512"], out["target512"] = n5, c5
    run_case(quiet(raletransformer.ralenet),
             synth_weights.make_state_dict("nra", 1, 2025), torch.from_numpy(n5), torch.from_numpy(c5),
             out, "nra512", evaluate)
    # case newrale: Transfer_learning.py:71-75
    n12, c12 = synth.make_batch(3, 12, 256, seed=99, kind="bw", snr_db=2.0)
    out["x12"], out["target12"] = n12, c12
    core = quiet(twelve.ralenet, high_level_enhence=True)
    run_case(twelve.newrale(core), synth_weights.make_newrale_state_dict(2023),
             torch.from_numpy(n12), torch.from_numpy(c12), out, "newrale", evaluate)

    path = os.path.join(a.out, "ralenet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
