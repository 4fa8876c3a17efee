"""12-lead transfer wrapper -- drop-in for the reference's `model/ralenet_12leads.py`
(which does not even import as shipped: it ends in a body-less `if __name__ == "__main__":`).

`ralenet` is the R-wave model with `transconv` registered right after `conv1` (reference :578-580);
`newrale` (reference :680-709) sandwiches a frozen, pretrained core between Conv1d(12->6,k13) ->
Conv1d(6->2,k13) and Conv1d(2->6,k13) -> Conv1d(6->12,k13) with LeakyReLU(0.01).
"""
from __future__ import annotations

import torch.nn as nn

from .. import ops
from ._blocks import (eca_layer_1d, AbsPositionalEncoding, BasicLayer, DropPath, LinearProjection, Mlp, MSAttention,  # noqa: F401
                      PartialConv_1d, PatchMerging, PatchSeparate, RelativePositionEmbedding, TransformerBlock,
                      _RalenetBase, build_ralenet, drop_path, mask_fill)


class ralenet(_RalenetBase):
    def __init__(self, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0., mlp_ratio=4., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, use_partial=True, use_eca=False, pe='abs', use_checkpoint=False,
                 low_level_enhence=True, high_level_enhence=False) -> None:
        super().__init__()
        build_ralenet(self, lambda c, h, le: BasicLayer(c, depth=2, num_heads=h, local_enhence=le), norm_layer,
                      high_level_enhence, with_rw=True, head_first=True)
        self._init_plan()


class newrale(nn.Module):
    def __init__(self, pretrained_rale_model):
        super(newrale, self).__init__()
        self.conv1 = nn.Conv1d(12, 6, kernel_size=13, padding=6)
        self.conv2 = nn.Conv1d(6, 2, kernel_size=13, padding=6)
        self.rale = pretrained_rale_model
        self.conv3 = nn.Conv1d(2, 6, kernel_size=13, padding=6)
        self.conv4 = nn.Conv1d(6, 12, kernel_size=13, padding=6)
        self.relu = nn.LeakyReLU()
        for param in self.rale.parameters():          # frozen core (reference :695-696)
            param.requires_grad = False

    def forward(self, x):
        s = self.relu.negative_slope
        x = ops.Conv1dFn.apply(x, self.conv1.weight, self.conv1.bias, True, s)
        x = ops.Conv1dFn.apply(x, self.conv2.weight, self.conv2.bias, True, s)
        x = self.rale(x)
        x = ops.Conv1dFn.apply(x, self.conv3.weight, self.conv3.bias, True, s)
        return ops.Conv1dFn.apply(x, self.conv4.weight, self.conv4.bias, False, s)
