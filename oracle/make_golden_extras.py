"""Generate tests/golden/extras_golden.npz by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_extras [--ref /root/reference] [--out tests/golden]

Pins the oracle's restatement of the two SURVEY.md section 8f extras:
  * eca_layer_1d and a TransformerBlock(use_eca=True) (model/transformer.py:100-113, 325-411), imported from the
    reference and run in float64 with autograd for the gradients;
  * single_snr_noise_add (local_utils/local_utils.py:176-192).  `local_utils.py` imports wfdb (absent here) at module
    level, so that ONE function is extracted from the reference source with `ast` and exec'd with numpy -- the code
    that runs is the reference's own text, nothing is copied into this repository.
"""
from __future__ import annotations

import argparse
import ast
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_function(path: str, name: str):
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    with contextlib.redirect_stdout(io.StringIO()):
        from model import transformer as T  # type: ignore
    out = {}
    g = torch.Generator().manual_seed(77)

    # ---- eca_layer_1d alone, k = 3 and 5
    for k in (3, 5):
        B, L, C = 3, 32, 64
        eca = T.eca_layer_1d(C, k_size=k).double()
        w = torch.randn(1, 1, k, generator=g, dtype=torch.float64) * 0.7
        eca.conv.weight.data.copy_(w)
        x = torch.randn(B, L, C, generator=g, dtype=torch.float64, requires_grad=True)
        gy = torch.randn(B, L, C, generator=g, dtype=torch.float64)
        y = eca(x)
        y.backward(gy)
        out[f"eca{k}/x"], out[f"eca{k}/w"], out[f"eca{k}/gy"] = x.detach().numpy(), w.numpy(), gy.numpy()
        out[f"eca{k}/y"], out[f"eca{k}/dx"] = y.detach().numpy(), x.grad.numpy()
        out[f"eca{k}/dw"] = eca.conv.weight.grad.numpy()

    # ---- a whole TransformerBlock with the gate between fc2 and the residual (stage s2: C = 32, H = 8, L = 64)
    C, H, L, B = 32, 8, 64, 2
    with contextlib.redirect_stdout(io.StringIO()):
        blk = T.TransformerBlock(C, H, local_enhence=True, use_eca=True).double()
    sd = blk.state_dict()
    for kname in sd:
        t = torch.randn(sd[kname].shape, generator=g, dtype=torch.float64)
        if kname.endswith("norm1.weight") or kname.endswith("norm2.weight"):
            t = 1 + 0.1 * t
        elif kname.endswith("bias"):
            t = 0.1 * t
        else:
            t = t * (sd[kname].shape[-1] ** -0.5)
        sd[kname] = t
    blk.load_state_dict(sd, strict=True)
    x = torch.randn(B, L, C, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, L, C, generator=g, dtype=torch.float64)
    y = blk(x)
    y.backward(gy)
    out["blk/x"], out["blk/gy"], out["blk/y"], out["blk/dx"] = x.detach().numpy(), gy.numpy(), y.detach().numpy(), x.grad.numpy()
    out["blk/keys"] = np.array(list(sd.keys()))
    for kname, p in blk.named_parameters():
        out[f"blk/p/{kname}"] = sd[kname].numpy()
        out[f"blk/g/{kname}"] = p.grad.numpy()

    # ---- single_snr_noise_add, per window, float32 as the reference's data pipeline runs it
    fn = reference_function(os.path.join(args.ref, "local_utils", "local_utils.py"), "single_snr_noise_add")
    rs = np.random.RandomState(5)
    data = rs.standard_normal((6, 256, 2)).astype(np.float32)          # (batch, length, channel), :194
    noise = (rs.standard_normal((6, 256, 2)) * rs.uniform(0.2, 3.0, (6, 1, 1))).astype(np.float32)
    snr = np.array([-4, -2, 0, 2, 4, 7.5], dtype=np.float32)
    out["snr/data"], out["snr/noise"], out["snr/snr"] = data, noise, snr
    out["snr/out"] = np.stack([fn(data[i], noise[i], float(snr[i])) for i in range(6)]).astype(np.float64)

    # ---- R-peak-positioned bias: RelativePositionEmbedding.forward(R_pos) (model/transformer.py:534-545) feeding a
    # TransformerBlock at stage s1 (C = 16, H = 4, L = 128, W = 16); R_pos = 40 puts the W x W block at offset 32
    C, H, L, B, W, R_POS = 16, 4, 128, 3, 16, 40
    with contextlib.redirect_stdout(io.StringIO()):
        blk = T.TransformerBlock(C, H, local_enhence=True).double()
    rw = T.RelativePositionEmbedding(W, L, H).double()
    rw.relative_position_bias_table.data = torch.randn(2 * W - 1, H, generator=g, dtype=torch.float64) * 0.5
    sd = blk.state_dict()
    for kname in sd:
        t = torch.randn(sd[kname].shape, generator=g, dtype=torch.float64)
        if kname.endswith("norm1.weight") or kname.endswith("norm2.weight"):
            t = 1 + 0.1 * t
        elif kname.endswith("bias"):
            t = 0.1 * t
        else:
            t = t * (sd[kname].shape[-1] ** -0.5)
        sd[kname] = t
    blk.load_state_dict(sd, strict=True)
    x = torch.randn(B, L, C, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, L, C, generator=g, dtype=torch.float64)
    for tag, rpos in (("rpos", R_POS), ("rpos_edge", W // 2)):          # W//2: the block touches the top-left corner
        blk.zero_grad(); rw.zero_grad(); x.grad = None
        y = blk(x, rw(rpos))
        y.backward(gy)
        out[f"{tag}/x"], out[f"{tag}/gy"], out[f"{tag}/y"], out[f"{tag}/dx"] = (x.detach().numpy(), gy.numpy(),
                                                                           y.detach().numpy(), x.grad.numpy().copy())
        out[f"{tag}/R_pos"] = np.int64(rpos)
        out[f"{tag}/table"] = rw.relative_position_bias_table.detach().numpy()
        out[f"{tag}/d_table"] = rw.relative_position_bias_table.grad.numpy().copy()
        out[f"{tag}/dense_mask"] = rw(rpos).detach().numpy()
        out[f"{tag}/keys"] = np.array(list(sd.keys()))
        for kname, p in blk.named_parameters():
            out[f"{tag}/p/{kname}"] = sd[kname].numpy()
            out[f"{tag}/g/{kname}"] = p.grad.numpy().copy()

    # ---- record -> windows: np_norm (local_utils/local_utils.py:261-266) per lead over the record, as
    # batch_norm_snr_iter applies it (:124, dim=0 of a (T, leads) signal), then the cut `signal[i:i+256, :]` for i in
    # the reference's own range expression (:53, `range(0, 650000, 256)`) with the record length substituted.
    lu = os.path.join(args.ref, "local_utils", "local_utils.py")
    np_norm = reference_function(lu, "np_norm")
    tree = ast.parse(open(lu).read())
    bdi = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "batch_data_iter")
    loop = next(n for n in ast.walk(bdi) if isinstance(n, ast.For) and isinstance(n.iter, ast.Call)
                and getattr(n.iter.func, "id", "") == "range")
    start, stop, step = (ast.literal_eval(a) for a in loop.iter.args)
    assert (start, stop, step) == (0, 650000, 256), (start, stop, step)
    sl = next(n for n in ast.walk(loop) if isinstance(n, ast.Subscript) and getattr(n.value, "id", "") == "signal")
    width = ast.literal_eval(sl.slice.elts[0].upper.right)              # signal[i:i + 256, :]
    assert width == 256
    rs = np.random.RandomState(9)
    T_rec = 5000
    rec = (rs.standard_normal((3, T_rec, 2)) * np.array([3.0, 0.4]) + np.array([1.5, -20.0])).astype(np.float32)
    wins = []
    for r in range(3):
        sig = np_norm(rec[r].astype(np.float64), dim=0)                 # (T, leads)
        for i in range(start, T_rec, step):
            w = sig[i:i + width, :]
            if w.shape[0] == width:                                     # full windows only (the reference's batching
                wins.append(w.T)                                        # never emits the ragged tail either)
    out["records/x"] = np.ascontiguousarray(rec.transpose(0, 2, 1))     # (R, leads, T), the layout of this repo
    out["records/windows"] = np.stack(wins)                             # (R*nper, leads, 256)
    out["records/cut"] = np.array([start, stop, step, width])

    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "extras_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.1f} KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
