"""Torch fp32 restatement of `Model.update` — the training step of the 2-D and 3-D RIFE models.  TEST INFRASTRUCTURE
(see oracle/__init__.py).

Reference:  Flow-3D/model/RIFE.py:81-275  (L1 student + L1 teacher + 0.1 * distillation, AdamW lr set per step)
            Flow-2D/model/RIFE.py:80-336  (LapLoss student/teacher, 0.01 * distillation with the NaN / > 10 guard, 1e-6 * |w|_1 of
                                           block2 + block_tea (detached: read through state_dict()), 1e-5 * photometric loss)
            Flow-2D/model/laplacian.py:10-75 (LapLoss, 5 levels, reflect-padded 5x5 Gaussian)
Pinned against the imported reference by tests/golden/make_update_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .ifnet_ref import IFNetRef


# ------------------------------------------------------------------------------------------------ Flow-2D/model/laplacian.py
def _gauss_kernel(channels, device):
    """laplacian.py:10-19: the 5x5 binomial kernel [1 4 6 4 1] (x) [1 4 6 4 1] / 256, one copy per channel (depth-wise)."""
    b = torch.tensor([1., 4., 6., 4., 1.])
    return (torch.outer(b, b) / 256.).repeat(channels, 1, 1, 1).to(device)


def _conv_gauss(img, kernel):
    """laplacian.py:33-36: reflect-pad by 2, depth-wise 5x5 convolution."""
    return F.conv2d(F.pad(img, (2, 2, 2, 2), mode="reflect"), kernel, groups=img.shape[1])


def _upsample(x):
    """laplacian.py:24-31 builds the zero-interleaved 2x image with two cat/view/permute round trips; the result is x at the even
    (row, column) positions and zeros elsewhere, then 4 * Gaussian."""
    n, c, h, w = x.shape
    up = x.new_zeros((n, c, 2 * h, 2 * w))
    up[:, :, ::2, ::2] = x
    return _conv_gauss(up, 4 * _gauss_kernel(c, x.device))


def _laplacian_pyramid(img, kernel, max_levels):
    """laplacian.py:38-57: level = current - upsample(downsample(blur(current))), cropped to the common size; next = the down-sample."""
    current, pyr = img, []
    for _ in range(max_levels):
        down = _conv_gauss(current, kernel)[:, :, ::2, ::2]
        up = _upsample(down)
        h, w = min(current.shape[2], up.shape[2]), min(current.shape[3], up.shape[3])
        pyr.append(current[:, :, :h, :w] - up[:, :, :h, :w])
        current = down
    return pyr


def lap_loss_ref(inp, target, max_levels=5):
    kernel = _gauss_kernel(1, inp.device)
    return sum(F.l1_loss(a, b) for a, b in zip(_laplacian_pyramid(inp, kernel, max_levels), _laplacian_pyramid(target, kernel, max_levels)))


# ------------------------------------------------------------------------------------------------ Flow-2D/model/RIFE.py:227-282
def _charbonnier(x, alpha=0.25, epsilon=1.e-9):
    return torch.pow(torch.pow(x, 2) + epsilon ** 2, alpha)


def _backward_warp(flow, frame):
    b, c, h, w = flow.size()
    frame = F.interpolate(frame, size=(h, w), mode="bilinear", align_corners=True)
    xx = torch.arange(0, w).view(1, -1).repeat(h, 1).view(1, 1, h, w).repeat(b, 1, 1, 1)
    yy = torch.arange(0, h).view(-1, 1).repeat(1, w).view(1, 1, h, w).repeat(b, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float().permute(0, 2, 3, 1).to(flow.device)
    grid = flow.permute(0, 2, 3, 1) + grid
    factor = torch.FloatTensor([[[[2 / w, 2 / h]]]]).to(flow.device)
    return F.grid_sample(frame, grid * factor - 1, align_corners=False)   # the reference relies on the default (False)


def _photometric(warped, frame1):
    h, w = warped.shape[2:]
    frame1 = F.interpolate(frame1, (h, w), mode="bilinear", align_corners=False)
    p = _charbonnier(warped - frame1)
    p = torch.sum(p, dim=1) / 3
    return torch.sum(p) / frame1.size(0)


# ------------------------------------------------------------------------------------------------ Model.update
class TrainerRef:
    """`Model` with `update` (the constructor's AdamW: Flow-3D/model/RIFE.py:29, Flow-2D/model/RIFE.py:26)."""

    def __init__(self, nd: int, flownet: IFNetRef | None = None):
        self.nd = nd
        self.flownet = flownet if flownet is not None else IFNetRef(nd)
        self.optimG = torch.optim.AdamW(self.flownet.parameters(), lr=1e-6, weight_decay=1e-3)

    def update(self, imgs, gt, learning_rate=0.0, training=True):
        """3-D: RIFE.py:81-275.  2-D: RIFE.py:80-336 on the `droplet2d` / `vimeo2d` dataset branch (1-channel frames, no flow gt)."""
        nd = self.nd
        for g in self.optimG.param_groups:
            g["lr"] = learning_rate
        img0, img1 = imgs[:, :1], imgs[:, 1:2]
        self.flownet.train(training)
        flow, mask, merged, flow_tea, merged_tea, loss_distill = self.flownet.forward_train(torch.cat((imgs, gt), 1), (4, 2, 1))
        if nd == 3:
            loss_l1 = F.l1_loss(merged[2], gt)
            loss_tea = F.l1_loss(merged_tea, gt)
            loss_G = loss_l1 * 1 + loss_tea * 1 + loss_distill * 0.1
            extra = {}
        else:
            loss_l1 = lap_loss_ref(merged[2], gt).mean()
            loss_tea = lap_loss_ref(merged_tea, gt).mean()
            sd = self.flownet.state_dict()
            l1_reg = 0.
            for k in sd:
                if "block2" in k or "block_tea" in k:
                    l1_reg = l1_reg + torch.norm(sd[k], 1)
            loss_photo = _photometric(_backward_warp(flow[2][:, 2:4], merged[2]), img0)
            loss_photo = loss_photo + _photometric(_backward_warp(flow[2][:, :2], merged[2]), img1)
            loss_photo = loss_photo / 2
            if math.isnan(loss_distill) or loss_distill > 10.:
                loss_distill = torch.tensor(0.)
            loss_G = loss_l1 * 1 + loss_tea * 1 + loss_distill * 0.01 + l1_reg * 1e-6 + loss_photo * 1e-5 + torch.tensor(0.) * 0
            extra = {"l1_reg": l1_reg * 1e-6, "loss_photo": loss_photo * 1e-5}
        if training:
            self.optimG.zero_grad()
            loss_G.backward()
            self.optimG.step()
        out = {"loss_l1": loss_l1, "loss_tea": loss_tea, "loss_distill": loss_distill * (0.01 if nd == 2 else 1), "loss_G": loss_G,
               "merged_tea": merged_tea, "flow": flow[2] if nd == 3 else flow[2][:, :2], "flow_tea": flow_tea,
               "mask": mask[2]}
        out.update(extra)
        return merged[2], out


def training_triplet(nd: int, n: int, size: int, seed: int = 1234, shift: int = 2):
    """Seeded synthetic (img0, img1, gt): a box of random 4^nd tiles on a zero canvas (the rectangle generators of
    Datasets/create_rectangle_2d.py:89-121 / create_data_3d.py:41-105 in miniature), img1 = the box moved by 2*shift voxels along
    the last axis and shift along the first, gt = half way.  Returns fp32 CPU tensors (n,1,*[size]*nd)."""
    g = torch.Generator().manual_seed(seed)
    sp = [size] * nd
    out = []
    b0, b1 = size // 4, size // 4 + size // 2
    tiles = torch.randint(30, 256, [n] + [size // 8] * nd, generator=g).float() / 255.0
    box = tiles
    for a in range(nd):
        box = box.repeat_interleave(4, dim=1 + a)
    for k in range(3):                              # k = 0 (img0), 2 (img1), 1 (gt)
        v = torch.zeros([n, 1] + sp)
        sl = [slice(None), 0] + [slice(b0, b1)] * nd
        sl[2] = slice(b0 + k * shift // 2 * 1, b1 + k * shift // 2 * 1)
        sl[-1] = slice(b0 + k * shift, b1 + k * shift)
        v[tuple(sl)] = box
        out.append(v)
    return out[0], out[2], out[1]
