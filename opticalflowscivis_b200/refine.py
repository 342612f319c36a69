"""Refinement nets `Contextnet` / `Unet` — Flow-2D/model/refine.py:9-84, Flow-3D/model/refine.py:9-82 — behind
`IFNet(refine=True)` (the reference's module-level switch, Flow-2D/model/IFNet.py:32,140-142,255-273; off upstream).

Parameter containers with the reference's `state_dict()` keys (`contextnet.conv1.conv1.0.weight` … `unet.conv.bias`); every
convolution runs on the tcgen05 engines of the IFBlocks (bf16 operands, fp32 accumulate, PReLU in the epilogue), the feature warps
on ofsv_warp{2,3}d_f32 (multi-channel).  Channel counts follow the reference files: 2-D 1 / 9 / 1 (adapted to 1-channel data), 3-D
3 / 17 / 3 (left at the upstream RGB sizes there, and therefore not reachable from the 3-D IFNet — kept only as modules).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C, ops
from .ifnet import _ConvParams, _PReLUParams, _conv, _pack_conv, _pack_convT, _rup

C = 16      # refine.py:30 (2-D) / :30 (3-D)


def run_tap_layer(lay, x, n, in_sp):
    """One tap-form layer on the bf16 engines: stride-1 layers on the stacked halo kernel when it holds them, the rest on the
    per-tap tcgen05 kernel.  x channels-last [N][D][H][W][Cin_s] (D = 1 in 2-D).  Returns (y, logical output dims)."""
    d, osp = lay.desc(n, in_sp, _C.BF16, has_residual=False)
    y = torch.empty((n,) + tuple(osp) + (lay.cout_s,), device=x.device, dtype=torch.float32 if lay.out_f32 else torch.bfloat16)
    if lay.in_stride == 1 and not getattr(lay, "no_halo", False):
        try:
            ops.conv(d, x, lay.w_halo, lay.bias, lay.prelu, None, y, "halo")
            return y, osp
        except NotImplementedError:
            lay.no_halo = True
    ops.conv(d, x, lay.w_tc, lay.bias, lay.prelu, None, y, "tc")
    return y, osp


def _to_cl(x, nd, cs):
    """fp32 (N,C,*sp) tensor — or a list of up to eight, concatenated along the channels — -> bf16 channels-last [N][D][H][W][cs]
    with zero-padded channels, one launch (ofsv_pack_nhwc_bf16)."""
    return ops.pack_nhwc(list(x) if isinstance(x, (list, tuple)) else [x], cs)


def _from_cl(y, c, nd):
    """channels-last [N][D][H][W][Cs] -> fp32 (N,c,*sp): bf16 activations through ofsv_unpack_nhwc_f32; the fp32 outputs of the
    linear heads (8 stored channels) are a strided view made contiguous."""
    if y.dtype == torch.bfloat16:
        return ops.unpack_nhwc(y, c, nd)
    if nd == 2:
        return y[:, 0, :, :, :c].permute(0, 3, 1, 2).contiguous()
    return y[..., :c].permute(0, 4, 1, 2, 3).contiguous()


def _conv_layers(m, prelu):
    """Tap-form layer(s) of one conv + PReLU: ofsv_conv_tc holds at most 128 output channels per launch, so wider layers (the
    256-channel `down3` stage of the Unet) are split along Cout into 128-channel launches whose outputs are concatenated."""
    import types
    out = []
    for lo in range(0, m.cout, 128):
        hi = min(m.cout, lo + 128)
        mm = types.SimpleNamespace(nd=m.nd, k=m.k, pad=m.pad, stride=m.stride, cout=hi - lo, weight=m.weight[lo:hi], bias=m.bias[lo:hi])
        pp = None if prelu is None else types.SimpleNamespace(weight=prelu.weight[lo:hi])
        out.append(_pack_conv(mm, pp))
    return out


def _run_split(lays, x, n, sp):
    ys = [run_tap_layer(lay, x, n, sp) for lay in lays]
    return (ys[0][0] if len(ys) == 1 else torch.cat([y for y, _ in ys], -1)), ys[0][1]


class _Packed:
    """Tap-form layers of a module, re-packed when a parameter changes (same keying as IFBlock.layers())."""

    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _layers(self):
        key = self._key()
        if getattr(self, "_packed_key", None) != key:
            self._packed, self._packed_key = self._pack(), key
        return self._packed


class Conv2(nn.Module, _Packed):
    """refine.py:21-29: conv(in, out, 3, stride, 1) -> conv(out, out, 3, 1, 1), each with PReLU."""

    def __init__(self, nd, in_planes, out_planes, stride=2):
        super().__init__()
        self.nd = nd
        self.conv1 = _conv(nd, in_planes, out_planes, 3, stride, 1)
        self.conv2 = _conv(nd, out_planes, out_planes, 3, 1, 1)

    def _pack(self):
        return [_conv_layers(self.conv1[0], self.conv1[1]), _conv_layers(self.conv2[0], self.conv2[1])]

    def run(self, x, n, sp):
        for lays in self._layers():
            x, sp = _run_split(lays, x, n, sp)
        return x, sp

    def forward(self, *_a, **_k):
        raise RuntimeError("Conv2 is driven by Contextnet / Unet (channels-last bf16 activations inside libofsv)")


class Contextnet(nn.Module):
    """refine.py:31-56 (2-D) / :31-54 (3-D): four stride-2 Conv2 stages; after each one the flow is halved (resize 0.5, * 0.5) and the
    stage's features are backward-warped by it.  Returns [f1, f2, f3, f4] as fp32 (N, C_i, *sp/2^i) like the reference."""

    def __init__(self, nd=2, in_planes=None):
        super().__init__()
        self.nd = nd
        cin = in_planes if in_planes is not None else (1 if nd == 2 else 3)
        self.conv1 = Conv2(nd, cin, C)
        self.conv2 = Conv2(nd, C, 2 * C)
        self.conv3 = Conv2(nd, 2 * C, 4 * C)
        self.conv4 = Conv2(nd, 4 * C, 8 * C)

    @torch.no_grad()
    def forward(self, x, flow):
        nd = self.nd
        x, flow = ops._cuda_f32(x, "x"), ops._cuda_f32(flow, "flow")
        if any(s % 16 for s in x.shape[2:]):
            raise NotImplementedError("Contextnet: spatial dims must be multiples of 16")
        n = x.shape[0]
        sp = ((1,) if nd == 2 else ()) + tuple(x.shape[2:])
        mode = "bilinear" if nd == 2 else "trilinear"
        warp = ops.warp2d if nd == 2 else ops.warp3d
        a = _to_cl(x, nd, 16)
        out = []
        for i, m in enumerate((self.conv1, self.conv2, self.conv3, self.conv4)):
            a, sp = m.run(a, n, sp)
            flow = F.interpolate(flow, scale_factor=0.5, mode=mode, align_corners=False, recompute_scale_factor=False) * 0.5
            out.append(warp(_from_cl(a, C << i, nd), flow.contiguous()))
        return out


class Unet(nn.Module, _Packed):
    """refine.py:58-84 (2-D) / :56-82 (3-D)."""

    def __init__(self, nd=2, in_planes=None, out_planes=None):
        super().__init__()
        self.nd = nd
        cin = in_planes if in_planes is not None else (9 if nd == 2 else 17)
        self.cout = out_planes if out_planes is not None else (1 if nd == 2 else 3)
        self.down0 = Conv2(nd, cin, 2 * C)
        self.down1 = Conv2(nd, 4 * C, 4 * C)
        self.down2 = Conv2(nd, 8 * C, 8 * C)
        self.down3 = Conv2(nd, 16 * C, 16 * C)
        for i, (a, b) in enumerate(((32 * C, 8 * C), (16 * C, 4 * C), (8 * C, 2 * C), (4 * C, C))):
            setattr(self, f"up{i}", nn.Sequential(_ConvParams(nd, a, b, 4, 2, 1, True), _PReLUParams(b)))
        self.conv = _ConvParams(nd, C, self.cout, 3, 1, 1)

    def _pack(self):
        ups = []
        for i in range(4):
            up = getattr(self, f"up{i}")
            ups.append(_pack_convT(self.nd, up[0].weight.detach().float(), up[0].bias.detach().float(), up[1].weight.detach().float(),
                                   _rup(up[0].cout, 16), False))
        last = _pack_conv(self.conv, None)
        last.out_f32, last.cout_s = True, 8
        return ups + [last]

    @torch.no_grad()
    def forward(self, img0, img1, warped_img0, warped_img1, mask, flow, c0, c1):
        nd = self.nd
        x = torch.cat((img0, img1, warped_img0, warped_img1, mask, flow), 1)
        x = ops._cuda_f32(x, "unet input")
        if any(s % 16 for s in x.shape[2:]):
            raise NotImplementedError("Unet: spatial dims must be multiples of 16")
        n = x.shape[0]
        sp = ((1,) if nd == 2 else ()) + tuple(x.shape[2:])
        cl = lambda t: _to_cl(t, nd, t.shape[1])                                     # noqa: E731  (context features: 16..128 channels)
        s0, sp0 = self.down0.run(_to_cl(x, nd, _rup(x.shape[1], 16)), n, sp)
        s1, sp1 = self.down1.run(torch.cat((s0, cl(c0[0]), cl(c1[0])), -1), n, sp0)
        s2, sp2 = self.down2.run(torch.cat((s1, cl(c0[1]), cl(c1[1])), -1), n, sp1)
        s3, sp3 = self.down3.run(torch.cat((s2, cl(c0[2]), cl(c1[2])), -1), n, sp2)
        L = self._layers()
        y, spy = run_tap_layer(L[0], torch.cat((s3, cl(c0[3]), cl(c1[3])), -1), n, sp3)
        y, spy = run_tap_layer(L[1], torch.cat((y, s2), -1), n, spy)
        y, spy = run_tap_layer(L[2], torch.cat((y, s1), -1), n, spy)
        y, spy = run_tap_layer(L[3], torch.cat((y, s0), -1), n, spy)
        y, _ = run_tap_layer(L[4], y, n, spy)
        return torch.sigmoid(_from_cl(y, self.cout, nd))


def refine_merged(net, img0, img1, warped_img0, warped_img1, mask, flow, merged2):
    """Flow-2D/model/IFNet.py:255-273: c0/c1 = contextnet(img, flow half); tmp = unet(...); res = tmp[:, :3] * 2 - 1;
    merged[2] = clamp(merged[2] + res, 0, 1)."""
    nd = net.nd
    c0 = net.contextnet(img0, flow[:, :nd])
    c1 = net.contextnet(img1, flow[:, nd:2 * nd])
    tmp = net.unet(img0, img1, warped_img0, warped_img1, mask, flow, c0, c1)
    res = tmp[:, :3] * 2 - 1
    return torch.clamp(merged2 + res, 0, 1)
