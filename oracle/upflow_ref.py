"""Test helpers for the UPFlow network row.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The oracle for `UPFlow_net.forward_2_frame_v3` is the reference ITSELF: tests/golden/make_upflow_net_golden.py imports the
unmodified UPFlow/model/upflow.py in the build container, gives it the deterministic weights below, and records its outputs in
tests/golden/upflow_net.npz; the product loads the same weights on the GPU box and must reproduce the record.  The operator-level
restatements (normalize_features, torch_warp, occlusion check) below are pinned against the reference by the same script.
"""
from __future__ import annotations

import hashlib

import torch
import torch.nn.functional as F


def deterministic_state(shapes: dict, seed: int = 1234, head_gain: float = 0.04) -> dict:
    """{name: tensor} for {name: shape}: N(0, 2 / fan_in) weights (the scale of initialize_msra, pwc_modules.py:53-70) and small
    non-zero biases, each from a generator seeded by (seed, name) — reproducible on any machine without storing 13 MB of weights.
    The two flow heads are scaled by head_gain: with untrained MSRA heads every level adds ~5 px of noise flow (25 px at the output
    of a 128 x 192 pair), a chaotic regime in which no two arithmetic orders agree; 0.01 gives flows of a few pixels (a level's residual is doubled by each of the five up-samplings that follow)."""
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:4], "little")
        g = torch.Generator().manual_seed(h)
        if len(shp) == 1:
            out[name] = torch.randn(shp, generator=g) * 0.05
            if name.startswith(("flow_estimators.conv_last", "context_networks.convs.6")):
                out[name] *= head_gain
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            out[name] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
            if name.startswith(("flow_estimators.conv_last", "context_networks.convs.6")):
                out[name] *= head_gain
    return out


def normalize_features_ref(a: torch.Tensor, b: torch.Tensor):
    """network_tools.normalize_features((a, b), normalize=True, center=True, moments_across_channels=False,
    moments_across_images=False) — UPFlow/model/upflow.py:95-138."""
    out = []
    for f in (a, b):
        mean = torch.mean(f, dim=[2, 3], keepdim=True)
        var = torch.var(f, dim=[2, 3], keepdim=True)
        out.append((f - mean) / torch.sqrt(var + 1e-16))
    return out


def torch_warp_ref(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    """tools.torch_warp — UPFlow/utils/tools.py:1317-1361 (grid_sample, zeros padding, default align_corners=False, no mask)."""
    b, c, h, w = x.shape
    xx = torch.arange(0, w, device=x.device).view(1, 1, 1, w).expand(b, 1, h, w)
    yy = torch.arange(0, h, device=x.device).view(1, 1, h, 1).expand(b, 1, h, w)
    vgrid = torch.cat((xx, yy), 1).float() + flo
    vx = 2.0 * vgrid[:, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * vgrid[:, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack([vx, vy], dim=3), padding_mode="zeros", align_corners=False)


def occ_check_ref(flow_fw, flow_bw, alpha_1=0.1, alpha_2=0.5, scale=1, obj_out_all="obj"):
    """tools.occ_check_model.__call__ with occ_type = 'for_back_check' — UPFlow/utils/tools.py:543-719: the forward-backward check
    (:592-630, sum_abs_or_squar = True) and, for obj_out_all = 'obj' (UPFlow_net's setting, upflow.py:299), pixels whose flow leaves
    the frame (:683-710) are NOT counted as occluded (:713-719)."""
    length = lambda x: torch.sum(torch.pow(x ** 2, 0.5), dim=1, keepdim=True)       # noqa: E731
    mag = length(flow_fw) + length(flow_bw)
    diff_fw = flow_fw + torch_warp_ref(flow_bw, flow_fw)
    diff_bw = flow_bw + torch_warp_ref(flow_fw, flow_bw)
    thresh = alpha_1 * mag + alpha_2 / scale
    occ = [(length(diff_fw) < thresh).float(), (length(diff_bw) < thresh).float()]
    if obj_out_all == "all":
        return occ[0], occ[1]
    out = []
    for o, fl in zip(occ, (flow_fw, flow_bw)):
        b, _, h, w = fl.shape
        px = torch.arange(w, device=fl.device).view(1, 1, 1, w).float() + fl[:, 0:1]
        py = torch.arange(h, device=fl.device).view(1, 1, h, 1).float() + fl[:, 1:2]
        inside = ((px <= w - 1) & (px >= 0) & (py <= h - 1) & (py >= 0)).float()
        out.append(((o == 1) | (inside == 0)).float())
    return out[0], out[1]


def smooth_pair(b: int, h: int, w: int, seed: int = 7, shift: int = 3):
    """Synthetic (im1, im2): smooth random texture and a copy translated by `shift` px, in the value range of the KITTI
    normalisation (UPFlow/dataset/kitti_dataset.py:98-101: about [-0.45, 0.59])."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand((b, 3, h + 16, w + 16), generator=g)
    base = F.avg_pool2d(base, 7, 1, 3)
    base = (base - base.mean()) / base.std() * 0.2
    return base[:, :, 8:-8, 8:-8].contiguous(), base[:, :, 8:-8, 8 - shift:-8 - shift].contiguous()
