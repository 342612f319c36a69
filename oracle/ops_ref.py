"""Torch fp32 restatement of the reference's L1 operators.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Every function cites the reference lines it follows.  The functions are device-agnostic:
on ``cpu`` they reproduce the reference's CPU path, on ``cuda`` its CUDA-eager path
(the two differ by ~1e-5 in the warp because ATen's CUDA division by a
python scalar is a multiplication by the reciprocal).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _lin(n: int, device) -> torch.Tensor:
    return torch.linspace(-1.0, 1.0, n, device=device)


def warp2d_ref(tenInput: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    """Flow-2D/model/warplayer.py:7-26 — base grid + flow/((S-1)/2) -> grid_sample(bilinear, border, align_corners)."""
    n, _, h, w = tenFlow.shape
    dev = tenFlow.device
    base = torch.stack(
        [_lin(w, dev).view(1, 1, w).expand(n, h, w), _lin(h, dev).view(1, h, 1).expand(n, h, w)], dim=1
    )                                                                     # :12-17
    fx = tenFlow[:, 0:1] / ((tenInput.shape[3] - 1.0) / 2.0)              # :19
    fy = tenFlow[:, 1:2] / ((tenInput.shape[2] - 1.0) / 2.0)              # :20
    g = (base + torch.cat([fx, fy], 1)).permute(0, 2, 3, 1)               # :25
    return F.grid_sample(tenInput, g, mode="bilinear", padding_mode="border", align_corners=True)  # :26


def warp3d_ref(tenInput: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    """Flow-3D/model/warplayer.py:9-41 — axis-rotating trilinear warp.

    Grid channel 0 is a linspace over dim 3, channel 1 over dim 2, channel 2 over dim 4 (:15-22),
    normalised by (shape[3]-1)/2, (shape[2]-1)/2, (shape[4]-1)/2 (:24-26); grid_sample reads them as
    (x -> dim 4, y -> dim 3, z -> dim 2).
    """
    n, _, d, h, w = tenFlow.shape
    dev = tenFlow.device
    g0 = _lin(h, dev).view(1, 1, h, 1).expand(n, d, h, w)
    g1 = _lin(d, dev).view(1, d, 1, 1).expand(n, d, h, w)
    g2 = _lin(w, dev).view(1, 1, 1, w).expand(n, d, h, w)
    base = torch.stack([g0, g1, g2], dim=1)
    f0 = tenFlow[:, 0:1] / ((tenInput.shape[3] - 1.0) / 2.0)
    f1 = tenFlow[:, 1:2] / ((tenInput.shape[2] - 1.0) / 2.0)
    f2 = tenFlow[:, 2:3] / ((tenInput.shape[4] - 1.0) / 2.0)
    g = (base + torch.cat([f0, f1, f2], 1)).permute(0, 2, 3, 4, 1)       # :31
    return F.grid_sample(tenInput, g, mode="bilinear", padding_mode="border", align_corners=True)  # :37


def blend_ref(w0: torch.Tensor, w1: torch.Tensor, mask_logit: torch.Tensor) -> torch.Tensor:
    """Flow-2D/model/IFNet.py:189,240 / Flow-3D/model/IFNet.py:186,242 — sigmoid mask blend."""
    m = torch.sigmoid(mask_logit)
    return w0 * m + w1 * (1 - m)


def corr81_ref(f1: torch.Tensor, f2: torch.Tensor, md: int = 4) -> torch.Tensor:
    """Cost volume of UPFlow/utils/pytorch_correlation.py:27-50 (kernel 1, pad = max_disp = 4, strides 1).

    out[b,(dy+md)*(2md+1)+(dx+md),y,x] = mean_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx], zero outside.
    Written as 81 shifted products instead of the reference's double unfold (same arithmetic,
    no (B,81,C,HW) blow-up); pinned against Corr_pyTorch in tests/golden/make_golden.py.
    """
    b, c, h, w = f1.shape
    f2p = F.pad(f2, (md, md, md, md))
    out = f1.new_empty(b, (2 * md + 1) ** 2, h, w)
    k = 0
    for dy in range(2 * md + 1):
        for dx in range(2 * md + 1):
            out[:, k] = (f1 * f2p[:, :, dy:dy + h, dx:dx + w]).mean(1)
            k += 1
    return out


def upsample2d_flow_as_ref(inputs: torch.Tensor, h: int, w: int, if_rate: bool = True) -> torch.Tensor:
    """UPFlow/model/pwc_modules.py:77-90 — bilinear align_corners=True resize, u*=w/w_, v*=h/h_."""
    res = F.interpolate(inputs, [h, w], mode="bilinear", align_corners=True)
    if if_rate:
        h_, w_ = inputs.shape[2:]
        res = torch.cat([res[:, 0:1] * (w / w_), res[:, 1:2] * (h / h_)], 1)
    return res


def warping_layer_no_div_ref(x: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """UPFlow/model/pwc_modules.py:184-207 — normalise by (S-1), sample with default align_corners=False,
    zeros padding, times the (grid_sample(ones) >= 1) validity mask."""
    b, c, h, w = x.shape
    xx = torch.arange(0, w, device=x.device).view(1, 1, 1, w).expand(b, 1, h, w)
    yy = torch.arange(0, h, device=x.device).view(1, 1, h, 1).expand(b, 1, h, w)
    vgrid = torch.cat((xx, yy), 1).float() + flow
    vx = 2.0 * vgrid[:, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * vgrid[:, 1] / max(h - 1, 1) - 1.0
    g = torch.stack([vx, vy], dim=3)
    out = F.grid_sample(x, g, padding_mode="zeros")
    mask = F.grid_sample(torch.ones_like(x), g)
    return out * (mask >= 1.0).float()


def resize_ref(x: torch.Tensor, scale_factor: float) -> torch.Tensor:
    """F.interpolate(..., scale_factor, bi/trilinear, align_corners=False) as used at
    Flow-2D/model/IFNet.py:89,92,115-116 and Flow-3D/model/IFNet.py:85,88,118-119."""
    mode = "bilinear" if x.dim() == 4 else "trilinear"
    return F.interpolate(x, scale_factor=scale_factor, mode=mode, align_corners=False, recompute_scale_factor=False)
