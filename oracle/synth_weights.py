"""Deterministic synthetic weights with the reference's state_dict layout.  TEST INFRASTRUCTURE ONLY.

`make_state_dict` restates the key order / shapes of the reference modules' `state_dict()`
(model/transformer.py:560-619, raletransformer.py:559-637, ralenet_12leads.py:562-696);
`oracle/make_golden.py` asserts, against the imported reference, that the layout is identical.
Values come from numpy's frozen legacy RandomState stream, so the golden fixtures do not need
to store the 1.09 M weights: tests regenerate them bit-identically from the seed.

Non-default values are chosen on purpose: R-wave bias tables are non-zero (the reference
initialises them to zero, which would hide bugs in the bias path), BatchNorm running stats are
perturbed, LayerNorm affine parameters are not (1, 0).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .ralenet_oracle import CHANNELS, HEADS, LAYERS, RW_WINDOW, block_prefix


def _block_entries(pre: str, C: int, le: int):
    e = [
        (pre + "attn.qkv_proj.to_q.weight", (C, C), "w", C),
        (pre + "attn.qkv_proj.to_q.bias", (C,), "b", 0),
        (pre + "attn.qkv_proj.to_kv.weight", (2 * C, C), "w", C),
        (pre + "attn.qkv_proj.to_kv.bias", (2 * C,), "b", 0),
        (pre + "attn.proj.weight", (C, C), "w", C),
        (pre + "attn.proj.bias", (C,), "b", 0),
        (pre + "norm1.weight", (C,), "g", 0), (pre + "norm1.bias", (C,), "b", 0),
        (pre + "norm2.weight", (C,), "g", 0), (pre + "norm2.bias", (C,), "b", 0),
        (pre + "mlp.fc1.weight", (4 * C, C), "w", C), (pre + "mlp.fc1.bias", (4 * C,), "b", 0),
        (pre + "mlp.fc2.weight", (C, 4 * C), "w", 4 * C), (pre + "mlp.fc2.bias", (C,), "b", 0),
    ]
    if le == 1:
        e.append((pre + "mlp.leconv.partial_conv3.weight", (1, 1, 3), "w", 3))
    elif le == 2:
        e.append((pre + "mlp.leconv.weight", (4 * C, 1, 3), "w", 3))
    return e


def layout(variant: str = "rw", le: int = 1):
    """[(key, shape, kind, fan_in)] in the reference's state_dict order.
    variant 'rw'  = model/transformer.py (and ralenet_12leads.ralenet, whose only difference is
                    that `transconv` is registered right after `conv1`: variant 'rw12');
    variant 'nra' = model/raletransformer.py."""
    ents = [
        ("conv1.0.weight", (8, 2, 3), "w", 6), ("conv1.0.bias", (8,), "b", 0),
        ("conv1.2.weight", (8,), "g", 0), ("conv1.2.bias", (8,), "b", 0),
        ("conv1.2.running_mean", (8,), "rm", 0), ("conv1.2.running_var", (8,), "rv", 0),
        ("conv1.2.num_batches_tracked", (), "nbt", 0),
    ]
    head = [("transconv.0.weight", (2, 8, 3), "w", 24), ("transconv.0.bias", (2,), "b", 0)]
    if variant == "rw12":
        ents += head
    base = "rw" if variant.startswith("rw") else "nra"
    if base == "rw":
        for i, W in enumerate(RW_WINDOW):
            ents.append((f"rwattn{i + 1}.relative_position_bias_table", (2 * W - 1, HEADS[i]), "t", 0))
            ents.append((f"rwattn{i + 1}.relative_position_index", (W, W), "idx", W))
    pm = {0: 1, 1: 2, 2: 3, 3: 4}
    ps = {5: 4, 6: 3, 7: 2, 8: 1}
    for li, (layer, s, _) in enumerate(LAYERS):
        C = CHANNELS[s]
        for i in range(2):
            ents += _block_entries(block_prefix(base, layer, i), C, le)
        if li in pm:
            j = pm[li]
            ents += [(f"pm{j}.reduction.weight", (2 * C, 2 * C), "w", 2 * C),
                     (f"pm{j}.norm.weight", (2 * C,), "g", 0), (f"pm{j}.norm.bias", (2 * C,), "b", 0)]
        if li in ps:
            j = ps[li]
            ents += [(f"ps{j}.reduction.weight", (C // 2, C // 2), "w", C // 2),
                     (f"ps{j}.norm.weight", (C // 2,), "g", 0), (f"ps{j}.norm.bias", (C // 2,), "b", 0)]
    if variant != "rw12":
        ents += head
    return ents


def make_state_dict(variant: str = "rw", le: int = 1, seed: int = 2023, dtype=torch.float32):
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for key, shape, kind, fan in layout(variant, le):
        if kind == "w":
            v = rs.standard_normal(shape) / np.sqrt(fan)
        elif kind == "b":
            v = 0.1 * rs.standard_normal(shape)
        elif kind == "g":
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif kind == "t":
            v = 0.5 * rs.standard_normal(shape)
        elif kind == "rm":
            v = 0.1 * rs.standard_normal(shape)
        elif kind == "rv":
            v = 0.5 + rs.uniform(size=shape)
        elif kind == "nbt":
            sd[key] = torch.tensor(3, dtype=torch.int64)
            continue
        elif kind == "idx":
            W = fan
            sd[key] = (torch.arange(W).view(-1, 1) - torch.arange(W).view(1, -1) + (W - 1)).to(torch.int64)
            continue
        sd[key] = torch.from_numpy(np.asarray(v, np.float64)).to(dtype)
    return sd


def make_newrale_state_dict(seed: int = 2023, dtype=torch.float32):
    """newrale (model/ralenet_12leads.py:681-696): conv1, conv2, rale.*, conv3, conv4."""
    rs = np.random.RandomState(seed + 7)
    sd = OrderedDict()

    def conv(name, co, ci):
        sd[name + ".weight"] = torch.from_numpy(rs.standard_normal((co, ci, 13)) / np.sqrt(13 * ci)).to(dtype)
        sd[name + ".bias"] = torch.from_numpy(0.1 * rs.standard_normal((co,))).to(dtype)

    conv("conv1", 6, 12)
    conv("conv2", 2, 6)
    for k, v in make_state_dict("rw12", 1, seed, dtype).items():
        sd["rale." + k] = v
    conv("conv3", 6, 2)
    conv("conv4", 12, 6)
    return sd
