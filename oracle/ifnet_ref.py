"""Torch fp32 restatement of IFBlock / IFNet / Model.inference, 2D and 3D.  TEST INFRASTRUCTURE
(see oracle/__init__.py).

One dimension-generic implementation; module attribute names are the reference's so that
``state_dict()`` keys are interchangeable (``block0.conv0.0.0.weight`` ... ``block_tea.conv2.2.bias``).
Pinned against the imported reference by tests/golden/make_golden.py (bit-identical outputs on CPU).

Reference:  Flow-2D/model/IFNet.py:16-27 (conv), :34-122 (IFBlock), :124-276 (IFNet)
            Flow-3D/model/IFNet.py:15-27, :31-120, :122-280
            Flow-2D/model/RIFE.py:66-78, Flow-3D/model/RIFE.py:67-79 (Model.inference)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .ops_ref import blend_ref, resize_ref, warp2d_ref, warp3d_ref


def _conv_prelu(nd, cin, cout, k, s, p):
    conv = (nn.Conv2d if nd == 2 else nn.Conv3d)(cin, cout, k, s, p, bias=True)
    return nn.Sequential(conv, nn.PReLU(cout))


def _head(nd, c, cout):
    ct = nn.ConvTranspose2d if nd == 2 else nn.ConvTranspose3d
    return nn.Sequential(ct(c, c // 2, 4, 2, 1), nn.PReLU(c // 2), ct(c // 2, cout, 4, 2, 1))


class IFBlockRef(nn.Module):
    """IFBlock, `version == 2` branch (the only one reachable: Flow-2D/model/IFNet.py:29, Flow-3D :29)."""

    def __init__(self, nd: int, in_planes: int, c: int):
        super().__init__()
        k0 = 3 if nd == 2 else 4                       # 3D conv0 is k=4,s=2,p=1
        self.nd = nd
        self.conv0 = nn.Sequential(_conv_prelu(nd, in_planes, c // 2, k0, 2, 1), _conv_prelu(nd, c // 2, c, k0, 2, 1))
        for i in range(4):
            setattr(self, f"convblock{i}", nn.Sequential(_conv_prelu(nd, c, c, 3, 1, 1), _conv_prelu(nd, c, c, 3, 1, 1)))
        self.conv1 = _head(nd, c, 2 * nd)              # flow: 4 ch (2D) / 6 ch (3D)
        self.conv2 = _head(nd, c, 1)                   # mask logit

    def forward(self, x, flow, scale):
        if scale != 1:
            x = resize_ref(x, 1.0 / scale)
        if flow is not None:
            flow = resize_ref(flow, 1.0 / scale) * 1.0 / scale
            x = torch.cat((x, flow), 1)
        x = self.conv0(x)
        for i in range(4):
            x = getattr(self, f"convblock{i}")(x) + x
        flow = resize_ref(self.conv1(x), scale) * scale
        mask = resize_ref(self.conv2(x), scale)
        return flow, mask


class IFNetRef(nn.Module):
    """3-scale student IFNet; inference branch only (gt has 0 channels).  Shapes must be multiples of 16
    so the reference's shape-repair slices (Flow-2D/model/IFNet.py:164-188) are no-ops."""

    WIDTHS = {2: (128, 96, 64), 3: (128, 64, 64)}      # Flow-2D/model/IFNet.py:127-129, Flow-3D :125-127

    def __init__(self, nd: int):
        super().__init__()
        self.nd = nd
        c0, c1, c2 = self.WIDTHS[nd]
        fl = 2 * nd
        self.block0 = IFBlockRef(nd, 2, c0)
        self.block1 = IFBlockRef(nd, 5 + fl, c1)
        self.block2 = IFBlockRef(nd, 5 + fl, c2)
        self.block_tea = IFBlockRef(nd, 6 + fl, 64)    # training only; kept for state_dict parity
        self.warp_fn = None                            # tests may inject another `warp(tenInput, tenFlow)` (e.g. the CUDA drop-in)

    def forward(self, x, scale=(4, 2, 1), timestep=0.5):   # timestep is ignored by the reference (fact 5)
        nd = self.nd
        warp = self.warp_fn or (warp2d_ref if nd == 2 else warp3d_ref)
        img0, img1 = x[:, :1], x[:, 1:2]
        flow_list, mask_list, merged = [], [], []
        w0, w1, flow, mask = img0, img1, None, None
        for i, blk in enumerate((self.block0, self.block1, self.block2)):
            if flow is None:
                flow, mask = blk(torch.cat((img0, img1), 1), None, scale[i])
            else:
                fd, md = blk(torch.cat((img0, img1, w0, w1, mask), 1), flow, scale[i])
                flow, mask = flow + fd, mask + md
            mask_list.append(torch.sigmoid(mask))
            flow_list.append(flow)
            w0 = warp(img0, flow[:, :nd])
            w1 = warp(img1, flow[:, nd:2 * nd])
            merged.append(w0 * mask_list[i] + w1 * (1 - mask_list[i]))
        return flow_list, mask_list, merged

    def forward_train(self, x, scale=(4, 2, 1)):
        """The `gt.shape[1] == 1` branch: student + teacher block + distillation loss.  x = cat(img0, img1, gt).
        Flow-3D/model/IFNet.py:133-280 (teacher :206-238, mask/loss :240-276); Flow-2D/model/IFNet.py:144-276 (teacher :206-230,
        distillation :238-248).  Returns the reference's tuple (flow_list, mask_list, merged, flow_teacher, merged_teacher,
        loss_distill) — the caller picks mask_list[2] in 3-D (IFNet.py:280)."""
        nd = self.nd
        warp = self.warp_fn or (warp2d_ref if nd == 2 else warp3d_ref)
        img0, img1, gt = x[:, :1], x[:, 1:2], x[:, 2:3]
        flow_list, mask_list, warped = [], [], []
        w0, w1, flow, mask = img0, img1, None, None
        for i, blk in enumerate((self.block0, self.block1, self.block2)):
            if flow is None:
                flow, mask = blk(torch.cat((img0, img1), 1), None, scale[i])
            else:
                fd, md = blk(torch.cat((img0, img1, w0, w1, mask), 1), flow, scale[i])
                flow, mask = flow + fd, mask + md
            mask_list.append(torch.sigmoid(mask))
            flow_list.append(flow)
            w0 = warp(img0, flow[:, :nd])
            w1 = warp(img1, flow[:, nd:2 * nd])
            warped.append((w0, w1))
        fd, md = self.block_tea(torch.cat((img0, img1, w0, w1, mask, gt), 1), flow, 1)
        flow_teacher = flow + fd
        w0t = warp(img0, flow_teacher[:, :nd])
        w1t = warp(img1, flow_teacher[:, nd:2 * nd])
        mask_teacher = torch.sigmoid(mask + md)
        merged_teacher = w0t * mask_teacher + w1t * (1 - mask_teacher)
        merged, loss_distill = [], 0
        for i in range(3):
            merged.append(warped[i][0] * mask_list[i] + warped[i][1] * (1 - mask_list[i]))
            loss_mask = ((merged[i] - gt).abs().mean(1, True) > (merged_teacher - gt).abs().mean(1, True) + 0.01).float().detach()
            loss_distill = loss_distill + (((flow_teacher.detach() - flow_list[i]) ** 2).mean(1, True) ** 0.5 * loss_mask).mean()
        return flow_list, mask_list, merged, flow_teacher, merged_teacher, loss_distill


class ModelRef:
    """Model.inference surface.  2D returns (merged[3], flow_list[3], mask_list[3]) (Flow-2D/model/RIFE.py:75);
    3D returns (merged[2], flow_list[3], mask_list[2]) (Flow-3D/model/RIFE.py:75, IFNet.py:280)."""

    def __init__(self, nd: int):
        self.nd = nd
        self.flownet = IFNetRef(nd)

    def eval(self):
        self.flownet.eval()
        return self

    @torch.no_grad()
    def inference(self, img0, img1, scale_list=(4, 2, 1), TTA=False, timestep=0.5):
        imgs = torch.cat((img0, img1), 1)
        flow, mask, merged = self.flownet(imgs, scale_list, timestep)
        if self.nd == 3:
            return merged[2], flow, mask[2]
        if not TTA:
            return merged, flow, mask
        _, _, merged2 = self.flownet(imgs.flip(2).flip(3), scale_list, timestep)   # Flow-2D/model/RIFE.py:77-78
        return (merged[2] + merged2[2].flip(2).flip(3)) / 2


__all__ = ["IFBlockRef", "IFNetRef", "ModelRef", "blend_ref"]
