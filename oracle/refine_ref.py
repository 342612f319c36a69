"""Torch fp32 restatement of the refinement nets Conv2 / Contextnet / Unet.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: Flow-2D/model/refine.py:9-84 (1 / 9 / 1 channels), Flow-3D/model/refine.py:9-82 (3 / 17 / 3 channels); use in
Flow-2D/model/IFNet.py:255-273 (`refine = True`).  Module attribute names are the reference's (state_dict keys interchange).
Pinned against the imported reference by tests/golden/make_refine_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .ops_ref import resize_ref, warp2d_ref, warp3d_ref

C = 16


def _conv(nd, cin, cout, k=3, s=1, p=1):
    return nn.Sequential((nn.Conv2d if nd == 2 else nn.Conv3d)(cin, cout, k, s, p, bias=True), nn.PReLU(cout))


def _deconv(nd, cin, cout):
    return nn.Sequential((nn.ConvTranspose2d if nd == 2 else nn.ConvTranspose3d)(cin, cout, 4, 2, 1, bias=True), nn.PReLU(cout))


class Conv2Ref(nn.Module):
    def __init__(self, nd, cin, cout, stride=2):
        super().__init__()
        self.conv1 = _conv(nd, cin, cout, 3, stride, 1)
        self.conv2 = _conv(nd, cout, cout, 3, 1, 1)

    def forward(self, x):
        return self.conv2(self.conv1(x))


class ContextnetRef(nn.Module):
    def __init__(self, nd=2):
        super().__init__()
        self.nd = nd
        self.conv1 = Conv2Ref(nd, 1 if nd == 2 else 3, C)
        self.conv2 = Conv2Ref(nd, C, 2 * C)
        self.conv3 = Conv2Ref(nd, 2 * C, 4 * C)
        self.conv4 = Conv2Ref(nd, 4 * C, 8 * C)

    def forward(self, x, flow):
        warp = warp2d_ref if self.nd == 2 else warp3d_ref
        out = []
        for m in (self.conv1, self.conv2, self.conv3, self.conv4):
            x = m(x)
            flow = resize_ref(flow, 0.5) * 0.5
            out.append(warp(x, flow))
        return out


class UnetRef(nn.Module):
    def __init__(self, nd=2):
        super().__init__()
        self.down0 = Conv2Ref(nd, 9 if nd == 2 else 17, 2 * C)
        self.down1 = Conv2Ref(nd, 4 * C, 4 * C)
        self.down2 = Conv2Ref(nd, 8 * C, 8 * C)
        self.down3 = Conv2Ref(nd, 16 * C, 16 * C)
        self.up0 = _deconv(nd, 32 * C, 8 * C)
        self.up1 = _deconv(nd, 16 * C, 4 * C)
        self.up2 = _deconv(nd, 8 * C, 2 * C)
        self.up3 = _deconv(nd, 4 * C, C)
        self.conv = (nn.Conv2d if nd == 2 else nn.Conv3d)(C, 1 if nd == 2 else 3, 3, 1, 1)

    def forward(self, img0, img1, warped_img0, warped_img1, mask, flow, c0, c1):
        s0 = self.down0(torch.cat((img0, img1, warped_img0, warped_img1, mask, flow), 1))
        s1 = self.down1(torch.cat((s0, c0[0], c1[0]), 1))
        s2 = self.down2(torch.cat((s1, c0[1], c1[1]), 1))
        s3 = self.down3(torch.cat((s2, c0[2], c1[2]), 1))
        x = self.up0(torch.cat((s3, c0[3], c1[3]), 1))
        x = self.up1(torch.cat((x, s2), 1))
        x = self.up2(torch.cat((x, s1), 1))
        x = self.up3(torch.cat((x, s0), 1))
        return torch.sigmoid(self.conv(x))


def refine_merged_ref(contextnet, unet, img0, img1, w0, w1, mask, flow, merged2, nd=2):
    """Flow-2D/model/IFNet.py:263-273."""
    c0 = contextnet(img0, flow[:, :nd])
    c1 = contextnet(img1, flow[:, nd:2 * nd])
    tmp = unet(img0, img1, w0, w1, mask, flow, c0, c1)
    return torch.clamp(merged2 + (tmp[:, :3] * 2 - 1), 0, 1)
