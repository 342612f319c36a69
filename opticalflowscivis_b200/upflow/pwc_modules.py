"""UPFlow/model/pwc_modules.py:77-90 (`upsample2d_flow_as`) and :179-207 (`WarpingLayer_no_div`)."""
from __future__ import annotations

import torch.nn as nn

from .. import ops


def upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate=False):
    if mode != "bilinear":
        raise NotImplementedError("only mode='bilinear' is used by the reference (upflow.py:608-609,622-623)")
    _, _, h, w = target_as.size()
    return ops.upsample_flow_ac(inputs, h, w, if_rate=if_rate)


class WarpingLayer_no_div(nn.Module):   # noqa: N801
    def forward(self, x, flow):
        return ops.warping_no_div(x, flow)
