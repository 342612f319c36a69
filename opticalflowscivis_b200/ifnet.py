"""Dimension-generic IFBlock / IFNet running on libofsv kernels.

Mirrors  Flow-2D/model/IFNet.py:34-122 (IFBlock), :124-276 (IFNet)  and  Flow-3D/model/IFNet.py:31-120, :122-280.
The modules below are PARAMETER CONTAINERS with the reference's `state_dict()` key names
(`block0.conv0.0.0.weight` ... `block_tea.conv2.2.bias`) and the reference's default initialisation; all arithmetic
runs in CUDA kernels behind the C ABI (no nn.Conv forward, no F.grid_sample, no F.interpolate).

Per block the launch sequence is (3-D, bf16 tensor-core engine)
    12 conv layers     conv0.{0,1} as stride-1 convs over the shifted space-to-depth input, convblock{0..3}.{0,1} (+residual),
                       conv1.0‖conv2.0 merged ConvT, conv1.2⊕conv2.2 as one depth-to-space conv whose epilogue accumulates the
                       flow/mask state when the block runs at scale 1
    block_stage_3d     F.interpolate(head, s)*s, flow += , mask += , sigmoid, warp x2, blend, and the NEXT block's resized concat
                       (bf16, space-to-depth) in one pass over the channels-last fp32 state
(block 0 starts from pack_block_input).  2-D and the fp32 validation engine use the unfused planar chain
    pack_block_input -> 12 conv layers -> head_upsample_add -> warp_blend.
"""
from __future__ import annotations

import itertools
import math

import torch
import torch.nn as nn

from . import _C, ops


USE_S2D = True     # stride-2 conv0 layers as stride-1 convs over the shifted space-to-depth input (halo engine)
USE_SHUFFLE_HEADS = True   # final ConvTranspose heads as one 3^d-tap depth-to-space conv (N = 2^d * 8) on the halo engine
USE_HALO = True    # stride-1 layers on the halo-reuse tcgen05 kernel (csrc/conv_halo.cu)
TC_READY = True    # the tcgen05 engine (csrc/conv_tc.cu) passed parity on B200 (tests/tc_probe.py, profiles/)


# ----------------------------------------------------------------------------------------------- parameter holders
class _ConvParams(nn.Module):
    """weight/bias of nn.Conv{2,3}d or nn.ConvTranspose{2,3}d with torch's default reset_parameters()."""

    def __init__(self, nd, cin, cout, k, stride, pad, transposed=False):
        super().__init__()
        self.nd, self.cin, self.cout, self.k, self.stride, self.pad, self.transposed = nd, cin, cout, k, stride, pad, transposed
        shape = ((cin, cout) if transposed else (cout, cin)) + (k,) * nd
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = self.weight.shape[1] * k ** nd
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, *_):
        raise RuntimeError("parameter container: the convolution runs inside libofsv (ofsv_conv_*)")


class _PReLUParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.full((c,), 0.25))

    def forward(self, *_):
        raise RuntimeError("parameter container: PReLU is fused into the conv epilogue")


def _conv(nd, cin, cout, k=3, s=1, p=1):
    return nn.Sequential(_ConvParams(nd, cin, cout, k, s, p), _PReLUParams(cout))


def _rup(x, m):
    return (x + m - 1) // m * m


# ----------------------------------------------------------------------------------------------- layer packing
class _Layer:
    """One conv layer in tap form + its packed device weights."""

    def __init__(self, nd, in_stride, out_stride, nphase, taps, w_tap, bias, prelu, cout_s, out_f32=False, residual=False,
                 shuffle=0):
        # taps: list (len nphase*ntaps) of (z,y,x) offsets; w_tap: fp32 [nphase*ntaps][Cin][Cout]
        self.nd, self.in_stride, self.out_stride, self.nphase = nd, in_stride, out_stride, nphase
        self.ntaps = len(taps) // nphase
        self.taps = taps
        cin, cout = w_tap.shape[1], w_tap.shape[2]
        self.cin_s, self.cout_w, self.cout_s = _rup(cin, 16), _rup(cout, 16), cout_s
        dev = w_tap.device
        w = torch.zeros(len(taps), self.cin_s, self.cout_w, device=dev, dtype=torch.float32)
        w[:, :cin, :cout] = w_tap
        self.w_simt = w.contiguous()
        self._packed = {}                  # weight layout -> bf16 blocks, packed by the library on first use (one launch)
        self.bias = torch.zeros(self.cout_w, device=dev, dtype=torch.float32)
        self.bias[:cout] = bias
        self.prelu = None
        if prelu is not None:
            self.prelu = torch.ones(self.cout_w, device=dev, dtype=torch.float32)
            self.prelu[:cout] = prelu
        self.out_f32, self.residual, self.shuffle = out_f32, residual, shuffle
        self.in_s2d = self.out_s2d = False
        self._desc_cache = {}

    def _structure_desc(self):
        """Descriptor carrying only what the weight layouts depend on (taps, Cin_s, Cout_w) — no shapes."""
        d = _C.ConvDesc()
        d.nd, d.Cin_s, d.Cout_w, d.Cout_s = self.nd, self.cin_s, self.cout_w, self.cout_s
        d.nphase, d.ntaps = self.nphase, self.ntaps
        for i, t in enumerate(self.taps):
            d.tap_off[i][0], d.tap_off[i][1], d.tap_off[i][2], d.tap_off[i][3] = t[0], t[1], t[2], 0
        return d

    def _pack(self, layout):
        w = self._packed.get(layout)
        if w is None:
            w = self._packed[layout] = ops.conv_pack_weights(self._structure_desc(), self.w_simt, layout)
        return w

    @property
    def w_tc(self):
        """bf16 [nphase][ntaps][Cin_s/KC][Cout_w][KC] blocks (ofsv_conv_tc, the plane-ring kernel)."""
        return self._pack(_C.WL_TAP)

    @property
    def w_halo(self):
        """The layout ofsv_conv_halo wants for this layer (stacked slots for the 3^d convs / ConvT / heads)."""
        return self._pack(ops.conv_halo_weight_layout(self._structure_desc()))

    def out_shape(self, n, osp):
        """Physical shape of the output tensor for logical output dims osp = (D,H,W)."""
        sp = osp if self.nd == 3 else osp[1:]
        if self.out_s2d:
            shp = [n] + [v // 2 + 1 for v in sp] + [(2 ** self.nd) * self.cout_s]
        else:
            shp = [n] + list(sp) + [self.cout_s]
        return shp

    def desc(self, n, in_sp, act_dtype, has_residual=None, hfast=False):
        """in_sp = LOGICAL (D,H,W) of the input (before space-to-depth); returns (ConvDesc, logical out_sp).  `has_residual`
        overrides the layer's own flag (state accumulation in the head conv), `hfast` selects the H-fastest depth-to-space output.
        Descriptors are cached per argument tuple (filling the tap table from Python costs ~40 us, more than the GPU time of a small
        layer) and must be treated as read-only."""
        key = (n, tuple(in_sp), act_dtype, has_residual, bool(hfast))
        hit = self._desc_cache.get(key)
        if hit is not None:
            return hit
        d = _C.ConvDesc()
        d.nd = self.nd
        d.N, (d.Di, d.Hi, d.Wi), d.Cin_s = n, in_sp, self.cin_s
        if self.in_s2d:
            d.Di, d.Hi, d.Wi = tuple((s // 2 + 1) if (self.nd == 3 or i > 0) else 1 for i, s in enumerate(in_sp))
            osp = tuple(max(1, s // 2) if (self.nd == 3 or i > 0) else 1 for i, s in enumerate(in_sp))
            vsp = osp
        elif self.shuffle:
            vsp = in_sp
            osp = tuple(s * 2 if (self.nd == 3 or i > 0) else 1 for i, s in enumerate(in_sp))
        elif self.nphase == 1:
            osp = tuple(max(1, s // self.in_stride) if (self.nd == 3 or i > 0) else 1 for i, s in enumerate(in_sp))
            vsp = osp
        else:
            vsp = in_sp
            osp = tuple(s * 2 if (self.nd == 3 or i > 0) else 1 for i, s in enumerate(in_sp))
        d.Do, d.Ho, d.Wo = vsp
        d.Dy, d.Hy, d.Wy = osp
        d.Cout_s, d.Cout_w = self.cout_s, self.cout_w
        d.in_stride, d.out_stride = self.in_stride, self.out_stride
        d.nphase, d.ntaps = self.nphase, self.ntaps
        for i, t in enumerate(self.taps):
            d.tap_off[i][0], d.tap_off[i][1], d.tap_off[i][2], d.tap_off[i][3] = t[0], t[1], t[2], 0
        d.has_prelu = int(self.prelu is not None)
        d.has_residual = int(self.residual if has_residual is None else has_residual)
        d.in_dtype = act_dtype
        d.out_dtype = _C.F32 if self.out_f32 else act_dtype
        d.out_shuffle = self.shuffle
        d.out_s2d = int(self.out_s2d)
        d.out_shuffle_hfast = int(bool(hfast))
        self._desc_cache[key] = (d, osp)
        return d, osp


def _conv_taps(nd, k, p):
    rng = [range(k)] * nd
    idx = list(itertools.product(*rng))
    offs = [((0,) if nd == 2 else ()) + tuple(i - p for i in ix) for ix in idx]
    return idx, offs


def _pack_conv(m: _ConvParams, prelu, residual=False):
    idx, offs = _conv_taps(m.nd, m.k, m.pad)
    w = m.weight.detach().float()                                    # [Cout][Cin][k..]
    w_tap = torch.stack([w[(slice(None), slice(None)) + ix].t() for ix in idx])   # [T][Cin][Cout]
    return _Layer(m.nd, m.stride, 1, 1, offs, w_tap, m.bias.detach().float(),
                  None if prelu is None else prelu.weight.detach().float(), _rup(m.cout, 16), residual=residual)


def _pack_conv_s2d(m: _ConvParams, prelu, cs_in, out_s2d):
    """Conv(k in {3,4}, stride 2, pad 1) as a stride-1 conv with tap offsets {0,1}^nd over the SHIFTED space-to-depth
    input (cell = (i+1)>>1, sub-cell = (i+1)&1 per axis; include/ofsv.h `out_s2d`): input i = 2o - 1 + k lives in cell
    o + (k>>1), sub-cell k&1, so kernel index k = 2*offset + sub.  cs_in = channel stride of one sub-cell."""
    nd, k = m.nd, m.k
    assert m.stride == 2 and m.pad == 1 and k in (3, 4) and not m.transposed
    nsub = 2 ** nd
    w = m.weight.detach().float()                                    # [Cout][Cin][k..]
    cout, cin = w.shape[0], w.shape[1]
    taps, w_tap = [], []
    for off in itertools.product((0, 1), repeat=nd):
        wt = torch.zeros(nsub * cs_in, cout, device=w.device)
        for sub in itertools.product((0, 1), repeat=nd):
            kk = tuple(2 * off[a] + sub[a] for a in range(nd))
            if all(v < k for v in kk):
                si = 0
                for a in range(nd):
                    si = (si << 1) | sub[a]                          # (z,y,x), x lowest bit — s2d_row() in ofsv_common.cuh
                wt[si * cs_in: si * cs_in + cin] = w[(slice(None), slice(None)) + kk].t()
        taps.append(((0,) if nd == 2 else ()) + off)
        w_tap.append(wt)
    lay = _Layer(nd, 1, 1, 1, taps, torch.stack(w_tap), m.bias.detach().float(),
                 None if prelu is None else prelu.weight.detach().float(), _rup(cout, 16))
    lay.in_s2d, lay.out_s2d = True, bool(out_s2d)
    return lay


def s2d_shift_pack(x, nd):
    """Channels-last [N][D][H][W][C] (D = 1 in 2-D) -> shifted space-to-depth [N][D/2+1][H/2+1][W/2+1][2^nd * C]
    (torch restatement of s2d_row(); used by the CPU tests)."""
    import torch.nn.functional as F
    n, d, h, w, c = x.shape
    if nd == 3:
        xp = F.pad(x, (0, 0, 1, 1, 1, 1, 1, 1))
        xp = xp.view(n, d // 2 + 1, 2, h // 2 + 1, 2, w // 2 + 1, 2, c).permute(0, 1, 3, 5, 2, 4, 6, 7)
        return xp.reshape(n, d // 2 + 1, h // 2 + 1, w // 2 + 1, 8 * c).contiguous()
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    xp = xp.view(n, 1, h // 2 + 1, 2, w // 2 + 1, 2, c).permute(0, 1, 2, 4, 3, 5, 6)
    return xp.reshape(n, 1, h // 2 + 1, w // 2 + 1, 4 * c).contiguous()


_CT_TAPS = {0: ((1, 0), (3, -1)), 1: ((2, 0), (0, 1))}   # ConvTranspose(4,2,1): parity -> ((kernel idx, input offset), ...)


def _pack_convT(nd, w, bias, prelu, cout_s, out_f32):
    """w [Cin][Cout][4..] (already merged / block-diagonal).  2^nd output parities x 2^nd taps."""
    taps, w_tap = [], []
    for par in itertools.product((0, 1), repeat=nd):                 # (z,y,x) parity; x lowest bit
        for choice in itertools.product((0, 1), repeat=nd):
            kk = tuple(_CT_TAPS[par[a]][choice[a]][0] for a in range(nd))
            off = tuple(_CT_TAPS[par[a]][choice[a]][1] for a in range(nd))
            taps.append(((0,) if nd == 2 else ()) + off)
            w_tap.append(w[(slice(None), slice(None)) + kk])         # [Cin][Cout]
    return _Layer(nd, 1, 2, 2 ** nd, taps, torch.stack(w_tap), bias, prelu, cout_s, out_f32=out_f32)


def _pack_heads_shuffle(nd, w, bias):
    """Depth-to-space form of the final ConvTranspose(4,2,1) heads (ofsv_conv_desc.out_shuffle = 8): ONE 3^nd-tap conv with
    2^nd * 8 output columns [output parity][8 channels]; a (parity, offset) pair that the transposed conv does not use
    gets zero weights.  w [Cin][Cout<=8][4..], bias [Cout]."""
    cin, cout = w.shape[0], w.shape[1]
    kidx = {(0, 0): 1, (0, -1): 3, (1, 0): 2, (1, 1): 0}            # (parity, input offset) -> kernel index, per axis
    taps, w_tap = [], []
    pars = list(itertools.product((0, 1), repeat=nd))
    for off in itertools.product((-1, 0, 1), repeat=nd):
        wt = torch.zeros(cin, 8 * len(pars), device=w.device)
        for pi, par in enumerate(pars):
            ks = [kidx.get((par[a], off[a])) for a in range(nd)]
            if all(k is not None for k in ks):
                wt[:, pi * 8: pi * 8 + cout] = w[(slice(None), slice(None)) + tuple(ks)]
        taps.append(((0,) if nd == 2 else ()) + off)
        w_tap.append(wt)
    b = torch.zeros(8 * len(pars), device=w.device)
    for pi in range(len(pars)):
        b[pi * 8: pi * 8 + cout] = bias
    return _Layer(nd, 1, 2, 1, taps, torch.stack(w_tap), b, None, 8, out_f32=True, shuffle=8)


class IFBlock(nn.Module):
    """Flow-2D/model/IFNet.py:34-122 / Flow-3D/model/IFNet.py:31-120, `version == 2` branch."""

    def __init__(self, nd, in_planes, c=64):
        super().__init__()
        self.nd, self.in_planes, self.c = nd, in_planes, c
        k0 = 3 if nd == 2 else 4
        self.conv0 = nn.Sequential(_conv(nd, in_planes, c // 2, k0, 2, 1), _conv(nd, c // 2, c, k0, 2, 1))
        for i in range(4):
            setattr(self, f"convblock{i}", nn.Sequential(_conv(nd, c, c), _conv(nd, c, c)))
        self.conv1 = nn.Sequential(_ConvParams(nd, c, c // 2, 4, 2, 1, True), _PReLUParams(c // 2),
                                   _ConvParams(nd, c // 2, 2 * nd, 4, 2, 1, True))
        self.conv2 = nn.Sequential(_ConvParams(nd, c, c // 2, 4, 2, 1, True), _PReLUParams(c // 2),
                                   _ConvParams(nd, c // 2, 1, 4, 2, 1, True))
        self._packed = None
        self._packed_key = None

    # -- weights -> tap form (re-done whenever a parameter changed or moved)
    def _key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def layers(self):
        key = self._key()
        if self._packed is None or self._packed_key != key:
            nd, c = self.nd, self.c
            L = [_pack_conv(self.conv0[0][0], self.conv0[0][1]), _pack_conv(self.conv0[1][0], self.conv0[1][1])]
            for i in range(4):
                cb = getattr(self, f"convblock{i}")
                L.append(_pack_conv(cb[0][0], cb[0][1]))
                L.append(_pack_conv(cb[1][0], cb[1][1], residual=True))
            # conv1.0 ‖ conv2.0: one ConvT c -> c (channels [0,c/2) feed the flow head, [c/2,c) the mask head)
            w10, w20 = self.conv1[0].weight.detach().float(), self.conv2[0].weight.detach().float()
            L.append(_pack_convT(nd, torch.cat([w10, w20], 1),
                                 torch.cat([self.conv1[0].bias, self.conv2[0].bias]).detach().float(),
                                 torch.cat([self.conv1[1].weight, self.conv2[1].weight]).detach().float(), c, False))
            # conv1.2 ⊕ conv2.2: block-diagonal ConvT c -> 2nd+1 (fp32 output, 8 stored channels)
            w12, w22 = self.conv1[2].weight.detach().float(), self.conv2[2].weight.detach().float()
            nf = 2 * nd
            wh = torch.zeros((c, nf + 1) + (4,) * nd, device=w12.device)
            wh[: c // 2, :nf] = w12
            wh[c // 2:, nf:] = w22
            bh = torch.cat([self.conv1[2].bias, self.conv2[2].bias]).detach().float()
            L.append(_pack_convT(nd, wh, bh, None, 8, True))
            self._heads_shuffle = _pack_heads_shuffle(nd, wh, bh)     # same layer, depth-to-space form (halo engine)
            # conv0.{0,1} in space-to-depth form (halo engine).  conv0.1's planes must fit shared memory: 2^nd*Cs0 channels
            # x (1 + dz range) planes <= 8 chunks of 64 channels -> always in 2-D, Cs0 <= 32 in 3-D
            cs0 = _rup(c // 2, 16)
            self._s2d1_ok = nd == 2 or cs0 <= 32
            self._s2d0 = _pack_conv_s2d(self.conv0[0][0], self.conv0[0][1], 16, out_s2d=self._s2d1_ok)
            self._s2d1 = _pack_conv_s2d(self.conv0[1][0], self.conv0[1][1], cs0, out_s2d=False) if self._s2d1_ok else None
            self._packed, self._packed_key = L, key
        return self._packed

    def can_accumulate_state(self, engine):
        """True when the final heads run as the depth-to-space conv whose epilogue can do `fm = fm_prev + head`."""
        return engine == "tc" and USE_HALO and USE_SHUFFLE_HEADS

    def run(self, xin, n, in_sp, act_dtype, engine, s2d_in=False, state_prev=None, hfast=False):
        """xin: packed channels-last block input [N][in_sp][16] (or its shifted space-to-depth form when s2d_in).
        Returns head [N][in_sp][8] fp32 — or, with `state_prev` ([N][in_sp][8] fp32 flow/mask state, scale-1 block only),
        the accumulated state state_prev + head written by the head conv's epilogue.  hfast (3-D depth-to-space heads only):
        head / state are H-fastest, [N][D][W][H][8] (ofsv_conv_desc.out_shuffle_hfast)."""
        L = self.layers()
        tdt = torch.float32 if act_dtype == _C.F32 else torch.bfloat16
        x, sp, skip = xin, in_sp, None
        for li, lay in enumerate(L):
            eng = engine(li, lay) if callable(engine) else engine
            if li == 11 and eng == "tc" and USE_HALO and USE_SHUFFLE_HEADS:
                lay = self._heads_shuffle
            if s2d_in and li == 0:
                lay = self._s2d0
            elif s2d_in and li == 1 and self._s2d1_ok:
                lay = self._s2d1
            acc = li == 11 and state_prev is not None
            if acc and not lay.shuffle:
                raise RuntimeError("state accumulation needs the depth-to-space head conv")
            if li == 11 and hfast and (not lay.shuffle or self.nd != 3):
                raise RuntimeError("the H-fastest state layout needs the 3-D depth-to-space head conv")
            d, osp = lay.desc(n, sp, act_dtype, has_residual=True if acc else None, hfast=(li == 11 and hfast))
            odt = torch.float32 if lay.out_f32 else tdt
            if li == 11 and hfast:
                y = torch.empty((n, osp[0], osp[2], osp[1], lay.cout_s), device=x.device, dtype=odt)
            elif lay.out_s2d:
                y = ops.workspace(("conv0", id(self)), lay.out_shape(n, osp), odt, x.device)
            else:
                y = torch.empty(lay.out_shape(n, osp), device=x.device, dtype=odt)
            res = skip if lay.residual else (state_prev if (li == 11 and state_prev is not None) else None)
            if eng == "tc" and USE_HALO and lay.in_stride == 1 and not getattr(lay, "no_halo", False):
                # stride-1 layers: halo-reuse kernel; layers it cannot hold in shared memory use the per-tap kernel
                try:
                    ops.conv(d, x, lay.w_halo, lay.bias, lay.prelu, res, y, "halo")
                    eng = None
                except NotImplementedError:
                    if lay.in_s2d:
                        raise
                    lay.no_halo = True
            if eng is not None:
                ops.conv(d, x, lay.w_simt if eng == "simt" else lay.w_tc, lay.bias, lay.prelu, res, y, eng)
            if 2 <= li <= 9 and (li % 2 == 0):
                skip = x                       # input of the residual pair
            x, sp = y, osp
        return x

    def forward(self, *_a, **_k):
        raise RuntimeError("IFBlock is driven by IFNet.forward (block input is built by ofsv_pack_block_input)")


class IFNet(nn.Module):
    """Flow-2D/model/IFNet.py:124-276 / Flow-3D/model/IFNet.py:122-280, inference branch (gt with 0 channels)."""

    WIDTHS = {2: (128, 96, 64), 3: (128, 64, 64)}

    def __init__(self, nd: int, precision: str = "bf16", engine: str = "auto", refine: bool = False):
        super().__init__()
        self.nd = nd
        self.refine = bool(refine)
        c0, c1, c2 = self.WIDTHS[nd]
        nf = 2 * nd
        self.block0 = IFBlock(nd, 2, c=c0)
        self.block1 = IFBlock(nd, 5 + nf, c=c1)
        self.block2 = IFBlock(nd, 5 + nf, c=c2)
        self.block_tea = IFBlock(nd, 6 + nf, c=64)      # teacher: training only (§8f), kept for state_dict parity
        if self.refine:                                 # Flow-2D/model/IFNet.py:140-142 (commented out in Flow-3D/model/IFNet.py:130-131)
            if nd != 2:
                raise NotImplementedError("refine=True exists for the 2-D IFNet only: Flow-3D/model/refine.py keeps the upstream RGB "
                                          "channel counts (3 / 17 / 3) and its use in Flow-3D/model/IFNet.py:274-279 is commented out")
            from .refine import Contextnet, Unet
            self.contextnet = Contextnet(nd)
            self.unet = Unet(nd)
        self.set_precision(precision, engine)
        self.only_last = False
        self.fuse_state_accumulate = True # scale-1 block: `flow += flow_d, mask += mask_d` inside the head conv's epilogue
        self.fuse_output_stage = True     # 3-D bf16: ofsv_block_stage_3d instead of head_upsample_add + warp_blend + pack
        self.state_hfast = True           # fused stage on the H-fastest state layout [N][D][W][H][8] (csrc/block_stage_hfast.cu)

    def set_precision(self, precision: str, engine: str = "auto"):
        """precision 'bf16' (tensor-core operands, fp32 accumulate; flow/mask accumulators and heads in fp32) or
        'fp32' (CUDA-core exact-order path).  engine 'tc' | 'simt' | 'auto' (tc wherever precision is bf16)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if engine not in ("auto", "tc", "simt"):
            raise ValueError("engine must be 'auto', 'tc' or 'simt'")
        if engine == "tc" and precision != "bf16":
            raise ValueError("the tensor-core engine computes in bf16")
        self.precision, self.engine = precision, engine

    def _engine(self):
        if self.engine == "auto":
            return "tc" if (self.precision == "bf16" and TC_READY) else "simt"
        return self.engine

    @torch.no_grad()
    def forward(self, x, scale=(4, 2, 1), timestep=0.5):
        """x = cat(img0, img1) (N,2,·) fp32 CUDA.  `timestep` is accepted and ignored exactly like the reference
        (SURVEY.md fact 5).  Returns (flow_list, mask_list | mask_list[2], merged, None, None, 0)."""
        nd = self.nd
        if x.dim() != nd + 2 or x.shape[1] < 2:
            raise ValueError(f"IFNet{nd}D: expected (N,2,{'D,' if nd == 3 else ''}H,W), got {tuple(x.shape)}")
        if x.shape[1] > 2:
            raise NotImplementedError("teacher/distillation branch (gt channel) is the training path, SURVEY.md §8f")
        return self.forward_pair(x[:, 0:1].contiguous(), x[:, 1:2].contiguous(), scale, timestep)

    @torch.no_grad()
    def forward_pair(self, img0, img1, scale=(4, 2, 1), timestep=0.5):
        """`forward` on the two frames/volumes given separately ((N,1,·) each): what `Model.inference` calls, so that the
        reference's `torch.cat((img0, img1), 1)` (RIFE.py:67 / :68) and the channel slicing that undoes it (IFNet.py:146-147)
        never touch HBM (1 GB of copies per 256^3 pair)."""
        nd = self.nd
        if img0.dim() != nd + 2 or img0.shape[1] != 1 or img1.shape != img0.shape:
            raise ValueError(f"IFNet{nd}D: expected two (N,1,{'D,' if nd == 3 else ''}H,W) tensors, got {tuple(img0.shape)} / {tuple(img1.shape)}")
        x = img0
        sp = tuple(x.shape[2:])
        if any(s % 16 for s in sp):
            raise NotImplementedError(f"spatial dims must be multiples of 16 (got {sp}); the reference's shape-repair "
                                      "slicing (Flow-2D/model/IFNet.py:164-188) is not reproduced")
        img0, img1 = ops._cuda_f32(img0, "img0"), ops._cuda_f32(img1, "img1")
        n = x.shape[0]
        act = _C.F32 if self.precision == "fp32" else _C.BF16
        eng = self._engine()
        flow_list, mask_list, merged = [], [], []
        w0 = w1 = flow = mask = None
        blocks = (self.block0, self.block1, self.block2)
        scales = [int(v) for v in scale]
        fused = nd == 3 and act == _C.BF16 and self.fuse_output_stage
        s2d = eng == "tc" and USE_S2D and USE_HALO and all((v // sc) % 4 == 0 for v in sp for sc in scales)
        xin = fm = None
        for i, blk in enumerate(blocks):
            s = scales[i]
            last = i == 2
            want_out = last or not self.only_last
            if xin is None:
                xin = ops.pack_block_input(img0, img1, w0, w1, mask, flow, s, act, s2d=s2d)
            in_sp = tuple(v // s for v in sp)
            # scale-1 block on the halo engine: the head conv's epilogue accumulates the state (fm = fm_prev + head)
            acc_state = fused and s == 1 and fm is not None and self.fuse_state_accumulate and blk.can_accumulate_state(eng)
            hfast = fused and self.state_hfast and blk.can_accumulate_state(eng)
            head = blk.run(xin, n, ((1,) + in_sp) if nd == 2 else in_sp, act, eng, s2d_in=s2d,
                           state_prev=fm if acc_state else None, hfast=hfast)
            xin = None
            if fused:
                # one pass over the channels-last state: resize + accumulate + warp x2 (+ blend) (+ the next block's input)
                s_next = 0 if last else (scales[i + 1] if scales[i + 1] in (1, 2) else 0)
                if acc_state:
                    fm, mg, ms, xin = ops.block_stage_3d(None, head, img0, img1, 0, s_next, want_out, want_out, pack_s2d=s2d, hfast=hfast)
                else:
                    fm, mg, ms, xin = ops.block_stage_3d(head, fm, img0, img1, s, s_next, want_out, want_out, pack_s2d=s2d, hfast=hfast)
                flow, mask = ops.state_views(fm, hfast)
                if not last and xin is None:          # next scale not fusable (4): fall back to the separate builder
                    flow, mask = flow.contiguous(), mask.contiguous()
                    w0, w1, _, _ = ops.warp_blend(img0, img1, flow, None, want_merged=False, want_mask=False)
            else:
                flow, mask = ops.head_upsample_add(head, flow, mask, nd, n, sp, s)
                w0, w1, mg, ms = ops.warp_blend(img0, img1, flow, mask, want_warped=True, want_merged=want_out, want_mask=want_out)
            flow_list.append(flow)
            mask_list.append(ms)
            merged.append(mg)
        if self.refine:
            if act != _C.BF16:
                raise NotImplementedError("the refinement nets run on the bf16 tensor-core engine only")
            from .refine import refine_merged
            merged[2] = refine_merged(self, img0, img1, w0, w1, mask, flow, merged[2])
        return flow_list, (mask_list if nd == 2 else mask_list[2]), merged, None, None, 0
