"""RA-LENet with R-wave attention -- drop-in for the reference's `model/transformer.py`.

Same public names (ralenet, BasicLayer, TransformerBlock, MSAttention, Mlp, PatchMerging, PatchSeparate,
RelativePositionEmbedding, mask_fill, ...), same constructor signatures and state_dict layout
(reference model/transformer.py:560-667); forward/backward run on hand-written sm_100a kernels.
"""
from __future__ import annotations

import torch.nn as nn

from ._blocks import (eca_layer_1d, AbsPositionalEncoding, BasicLayer, DropPath, LinearProjection, Mlp, MSAttention,  # noqa: F401
                      PartialConv_1d, PatchMerging, PatchSeparate, RelativePositionEmbedding, TransformerBlock,
                      _RalenetBase, build_ralenet, drop_path, mask_fill)


class ralenet(_RalenetBase):
    """reference model/transformer.py:560-667.  As in the reference only `high_level_enhence` and `norm_layer`
    are consumed; every other keyword is accepted and ignored (`low_level_enhence` included)."""

    def __init__(self, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0., mlp_ratio=4., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, use_partial=True, use_eca=False, pe='abs', use_checkpoint=False,
                 low_level_enhence=True, high_level_enhence=False) -> None:
        super().__init__()
        build_ralenet(self, lambda c, h, le: BasicLayer(c, depth=2, num_heads=h, local_enhence=le), norm_layer,
                      high_level_enhence, with_rw=True, head_first=False)
        self._init_plan()
