"""RA-LENet without R-wave attention ("ralenet_nra") -- drop-in for the reference's `model/raletransformer.py`.

Differences from `transformer.py` that are part of the contract (reference model/raletransformer.py:559-680):
local enhancement defaults to ON in Mlp / TransformerBlock / BasicLayer, the nine layers are
`nn.Sequential`s (state_dict keys `dtransformer1.0....`), there are no `rwattn*` tables, `BasicLayer` takes
channels-first input, and windows of any length with L % 16 == 0 work (the kernels support 256 and 512).
"""
from __future__ import annotations

import torch.nn as nn

from . import _blocks as _b
from ._blocks import (eca_layer_1d, AbsPositionalEncoding, DropPath, LinearProjection, MSAttention, PartialConv_1d,  # noqa: F401
                      PatchMerging, PatchSeparate, _RalenetBase, build_ralenet, drop_path)


class Mlp(_b.Mlp):
    _LE_DEFAULT = True


class TransformerBlock(_b.TransformerBlock):
    _LE_DEFAULT = True
    _MLP = Mlp


class BasicLayer(_b.BasicLayer):
    _LE_DEFAULT = True
    _BLOCK = TransformerBlock
    _CHANNELS_FIRST = True

    def forward(self, x):
        return super().forward(x, None)


class ralenet(_RalenetBase):
    """reference model/raletransformer.py:559-680: no keyword except `norm_layer` is consumed."""

    def __init__(self, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0., mlp_ratio=4., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, local_enhence=False, use_partial=True, use_eca=False, pe='abs',
                 use_checkpoint=False) -> None:
        super().__init__()
        build_ralenet(self, lambda c, h, le: nn.Sequential(TransformerBlock(c, num_heads=h),
                                                           TransformerBlock(c, num_heads=h)),
                      norm_layer, True, with_rw=False, head_first=False)
        self._init_plan()
