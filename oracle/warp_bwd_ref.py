"""numpy restatement of the BACKWARD of the reference warp.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows ATen's grid_sampler_{2,3}d_backward (bilinear, padding_mode=border, align_corners=True) as autograd runs it under
Flow-2D/model/warplayer.py:26 and Flow-3D/model/warplayer.py:37, chained with the backward of the `tenFlow / ((S-1)/2)`
normalisation (:19-20 / :24-26).  Pinned by tests/golden/make_warp_bwd_golden.py against autograd through the imported
reference (tests/golden/warp_bwd.npz); coordinates are evaluated in fp32 with the forward oracle's operation order, the
gradient sums in float64 (so it is an independent check of the kernels' summation, not a bit-level twin).
"""
from __future__ import annotations

import numpy as np
import torch

f32 = np.float32


def _lin(n):
    return torch.linspace(-1.0, 1.0, n).numpy()


def _coords(flow, sizes_sampled, half_extents, lins):
    """Per flow channel a: un-normalised, border-clipped coordinate along the axis it samples + clip gradient."""
    out = []
    for a, (S, he, lin) in enumerate(zip(sizes_sampled, half_extents, lins)):
        g = (lin + flow[:, a] / f32(he)).astype(f32)
        u = (((g + f32(1)) * f32(0.5)).astype(f32) * f32(S - 1)).astype(f32)
        clipgrad = ((u > 0) & (u < S - 1)).astype(np.float64)          # clip_coordinates_set_grad
        i = np.minimum(f32(S - 1), np.maximum(u, f32(0))).astype(f32)
        out.append((i, clipgrad))
    return out


def warp_bwd(src: np.ndarray, flow: np.ndarray, gout: np.ndarray):
    """Returns (grad_src, grad_flow) as float32 arrays; src/gout (N,C,*sp), flow (N,nd,*sp), nd = 2 or 3."""
    nd = flow.shape[1]
    N, C = src.shape[:2]
    sp = src.shape[2:]
    if nd == 2:
        H, W = sp
        lins = [_lin(W).reshape(1, 1, W), _lin(H).reshape(1, H, 1)]
        sampled = [W, H]                      # channel 0 -> x (W axis), 1 -> y (H axis)
        halves = [(W - 1.0) / 2.0, (H - 1.0) / 2.0]
        strides = [1, W]
    else:
        D, H, W = sp
        # Flow-3D/model/warplayer.py:15-26: channel 0 = linspace over dim 3 (H entries) normalised by (H-1)/2 but read by
        # grid_sample as x (W axis); channel 1 over dim 2 (D) / (D-1)/2 read as y (H axis); channel 2 over dim 4 (W) read as z
        lins = [_lin(H).reshape(1, 1, H, 1), _lin(D).reshape(1, D, 1, 1), _lin(W).reshape(1, 1, 1, W)]
        sampled = [W, H, D]
        halves = [(H - 1.0) / 2.0, (D - 1.0) / 2.0, (W - 1.0) / 2.0]
        strides = [1, W, H * W]
    co = _coords(flow.astype(f32), sampled, halves, lins)
    i0 = [np.floor(c[0]).astype(np.int64) for c in co]
    w1 = [(c[0] - np.floor(c[0])).astype(np.float64) for c in co]
    w0 = [1.0 - w for w in w1]
    inb = [(i0[a] + 1 <= sampled[a] - 1) for a in range(nd)]           # +1 neighbour inside the volume
    V = int(np.prod(sp))
    gsrc = np.zeros((N, C, V), np.float64)
    gi = [np.zeros((N,) + tuple(sp), np.float64) for _ in range(nd)]
    srcf = src.reshape(N, C, V).astype(np.float64)
    go = gout.astype(np.float64)
    nidx = np.arange(N).reshape((N,) + (1,) * nd)
    for corner in range(1 << nd):
        bits = [(corner >> a) & 1 for a in range(nd)]
        ok = np.ones((N,) + tuple(sp), bool)
        off = np.zeros((N,) + tuple(sp), np.int64)
        wgt = np.ones((N,) + tuple(sp), np.float64)
        for a in range(nd):
            off += (i0[a] + bits[a]) * strides[a]
            wgt = wgt * (w1[a] if bits[a] else w0[a])
            if bits[a]:
                ok &= inb[a]
        off = np.where(ok, off, 0)
        for c in range(C):
            val = np.where(ok, srcf[nidx, c, off], 0.0)                # ATen skips taps outside the volume
            np.add.at(gsrc[:, c], (np.broadcast_to(nidx, off.shape)[ok], off[ok]), (wgt * go[:, c])[ok])
            for a in range(nd):
                other = np.ones_like(wgt)
                for b in range(nd):
                    if b != a:
                        other = other * (w1[b] if bits[b] else w0[b])
                gi[a] += (1.0 if bits[a] else -1.0) * val * other * go[:, c]
    gflow = np.zeros(flow.shape, np.float64)
    for a in range(nd):
        gflow[:, a] = gi[a] * co[a][1] * ((sampled[a] - 1.0) / 2.0) / halves[a]
    return gsrc.reshape(src.shape).astype(f32), gflow.astype(f32)
