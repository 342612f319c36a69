"""Training step of the RIFE models on libofsv: `Model.update` — Flow-3D/model/RIFE.py:81-275,
Flow-2D/model/RIFE.py:80-336 — and the gt branch of `IFNet.forward` (teacher block + distillation loss:
Flow-3D/model/IFNet.py:206-276, Flow-2D/model/IFNet.py:206-248).

What runs where
  * the 14 conv / ConvTranspose / PReLU layers of every IFBlock, forward AND backward, run in libofsv: `_BlockFn` is ONE autograd
    node per block.  Forward = the tensor-core engines of inference (bf16 operands, fp32 accumulate) with every activation kept;
    backward per layer = ofsv_prelu_bias_bwd_bf16 (gradient through bias + PReLU, d bias, d slope), ofsv_conv_wgrad_bf16 (weight
    gradient) and the INPUT gradient as another tap-form layer on the same forward engines: the input gradient of a Conv(k, s=1)
    is the conv with mirrored taps and transposed weights, of a Conv(k, s=2, p=1) the ConvTranspose(k, 2, 1) phase form, of a
    ConvTranspose(4, 2, 1) the Conv(4, s=2, p=1) — all with the layer's own weights;
  * warp forward / backward: ofsv_warp{2,3}d_f32 / ofsv_warp{2,3}d_bwd_f32 behind `ops.warp2d / warp3d` (autograd Functions);
  * optimizer: `optim.FusedAdamW` on a flat `GradientBucket`, one all-reduce of the bucket when a process group is up (the
    reference wraps the net in DDP when local_rank != -1: RIFE.py:31-32);
  * the student blocks' input packing (resize 1/s of the concatenation, flow / s) and head up-sampling + accumulate are libofsv
    autograd nodes as well (ofsv_pack_block_input[_bwd], ofsv_head_upsample_add[_bwd]; the teacher's concat: ofsv_pack_nhwc_bf16 /
    ofsv_unpack_nhwc_f32); what is left between the nodes (sigmoid, blend, the loss reductions) is plain torch autograd on fp32
    NC(D)HW tensors — element-wise launches, none of them a convolution or a sampler.
There is no CPU path: everything raises on CPU tensors like the rest of the package.
"""
from __future__ import annotations

import ctypes
import itertools
import math

import torch
import torch.nn.functional as F

from . import _C, ops
from .ifnet import IFBlock, IFNet, _Layer, _conv_taps, _rup

# ConvTranspose(k, 2, 1) per axis: output parity -> ((kernel index | None, input offset), (…)); None = a tap the kernel does not have
_CT_TAPS = {4: {0: ((1, 0), (3, -1)), 1: ((2, 0), (0, 1))},
            3: {0: ((1, 0), (None, -1)), 1: ((2, 0), (0, 1))}}


def _convT_phase_layer(nd, w, k, cout_s):
    """Phase form (include/ofsv.h: nphase = ntaps = 2^nd, out_stride 2) of ConvTranspose(k in {3,4}, 2, 1[, output_padding = k == 3])
    with weight w [Cin][Cout][k..], no bias / activation, bf16 output: the input gradient of Conv(k, 2, 1) whose weight tensor
    [Cout_c][Cin_c][k..] is exactly this layout."""
    taps, w_tap = [], []
    zero = torch.zeros(w.shape[0], w.shape[1], device=w.device)
    for par in itertools.product((0, 1), repeat=nd):
        for choice in itertools.product((0, 1), repeat=nd):
            ks = [_CT_TAPS[k][par[a]][choice[a]][0] for a in range(nd)]
            off = tuple(_CT_TAPS[k][par[a]][choice[a]][1] for a in range(nd))
            taps.append(((0,) if nd == 2 else ()) + off)
            w_tap.append(zero if any(v is None for v in ks) else w[(slice(None), slice(None)) + tuple(ks)])
    cout = w.shape[1]
    return _Layer(nd, 1, 2, 2 ** nd, taps, torch.stack(w_tap), torch.zeros(cout, device=w.device), None, cout_s)


def _conv_layer(nd, w, k, stride, pad, cout_s, mirror=False):
    """Conv(k, stride, pad) with weight w [Cout][Cin][k..] as a tap-form layer without bias / activation.  mirror = True: the
    INPUT GRADIENT of the stride-1 conv with that weight instead (taps negated, per-tap matrices transposed: Cout -> Cin)."""
    idx, offs = _conv_taps(nd, k, pad)
    if mirror:
        assert stride == 1
        offs = [tuple(-v for v in o) for o in offs]
        w_tap = torch.stack([w[(slice(None), slice(None)) + ix] for ix in idx])          # [T][Cout][Cin]: rows = gradient channels
        cout = w.shape[1]
    else:
        w_tap = torch.stack([w[(slice(None), slice(None)) + ix].t() for ix in idx])      # [T][Cin][Cout]
        cout = w.shape[0]
    return _Layer(nd, stride, 1, 1, offs, w_tap, torch.zeros(cout, device=w.device), None, cout_s)


def _convT_kflat(nd):
    """Flat kernel index (C order over [4]*nd) of tap i of ifnet._pack_convT's (parity, choice) enumeration."""
    out = []
    for par in itertools.product((0, 1), repeat=nd):
        for choice in itertools.product((0, 1), repeat=nd):
            f = 0
            for a in range(nd):
                f = f * 4 + _CT_TAPS[4][par[a]][choice[a]][0]
            out.append(f)
    return out


# ------------------------------------------------------------------------------------------------ low-level wrappers
def _require_cuda(t, name):
    """No CPU path.  (tests/test_train_host.py swaps this check and the four kernel wrappers below for torch evaluators of the
    same tap forms, to verify the backward WIRING of this module in fp32 without a GPU.)"""
    if not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (no CPU path)")
    return t


_ACT_DTYPE = torch.bfloat16      # activation / gradient storage type of the conv stack


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def prelu_bias_bwd(gy, y, slope):
    """gy, y bf16 [..., Cs] channels-last; slope fp32 [Cs] or None.  Returns (gpre bf16 like gy, dbias fp32 [Cs], dslope | None)."""
    if not (gy.is_cuda and gy.dtype == torch.bfloat16 and gy.is_contiguous()):
        raise TypeError("prelu_bias_bwd: gy must be a contiguous CUDA bf16 tensor (no CPU path)")
    cs = gy.shape[-1]
    rows = gy.numel() // cs
    L = _C.lib()
    with ops._on(gy.device):
        nblk = L.ofsv_prelu_bias_bwd_blocks()
        work = torch.empty(2 * cs * nblk, device=gy.device, dtype=torch.float32)
        dbias = torch.empty(cs, device=gy.device, dtype=torch.float32)
        dslope = torch.empty(cs, device=gy.device, dtype=torch.float32) if slope is not None else None
        gpre = torch.empty_like(gy) if slope is not None else gy
        with ops._span("prelu_bias_bwd"):
            _C.check(L.ofsv_prelu_bias_bwd_bf16(_p(gy), _p(y), _p(slope), _p(gpre), _p(dbias), _p(dslope), _p(work), rows, cs, _stream()))
    return gpre, dbias, dslope


def conv_wgrad(desc, x, gy):
    """Weight gradient in tap form, fp32 [nphase*ntaps][Cin_s][Cout_w] (ofsv_conv_wgrad_bf16)."""
    for t, name in ((x, "x"), (gy, "gy")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
            raise TypeError(f"conv_wgrad: {name} must be a contiguous CUDA bf16 tensor (no CPU path)")
    L = _C.lib()
    T = desc.nphase * desc.ntaps
    n = T * desc.Cin_s * desc.Cout_w
    with ops._on(x.device):
        splits = L.ofsv_conv_wgrad_splits(ctypes.byref(desc))
        if splits < 0:
            _C.check(splits)
        dw = torch.empty(T, desc.Cin_s, desc.Cout_w, device=x.device, dtype=torch.float32)
        work = torch.empty(splits * n, device=x.device, dtype=torch.float32) if splits > 1 else None
        with ops._span("conv_wgrad"):
            _C.check(L.ofsv_conv_wgrad_bf16(ctypes.byref(desc), _p(x), _p(gy), gy.shape[-1], _p(dw), _p(work), _stream()))
    return dw


def _run_layer(lay, d, x, res, y):
    """Stride-1 layers on the stacked halo engine when it holds them, everything else on the per-tap tcgen05 engine."""
    if lay.in_stride == 1 and not getattr(lay, "no_halo", False):
        try:
            return ops.conv(d, x, lay.w_halo, lay.bias, lay.prelu, res, y, "halo")
        except NotImplementedError:
            lay.no_halo = True
    return ops.conv(d, x, lay.w_tc, lay.bias, lay.prelu, res, y, "tc")


def _to_cl16(x, nd):
    """fp32 (N,C,*sp) -> bf16 channels-last [N][D][H][W][16] (D = 1 in 2-D), zero-padded channels (ofsv_pack_nhwc_bf16)."""
    return ops.pack_nhwc([x], 16)


def _from_cl(y, c, nd):
    """channels-last [N][D][H][W][Cs] bf16 / fp32 -> fp32 (N,c,*sp) contiguous."""
    if y.dtype == torch.bfloat16:
        return ops.unpack_nhwc(y, c, nd)
    if nd == 2:                                    # the fp32 head tensor (8 stored channels)
        return y[:, 0, :, :, :c].permute(0, 3, 1, 2).contiguous()
    return y[..., :c].permute(0, 4, 1, 2, 3).contiguous()


# ------------------------------------------------------------------------------------------------ one IFBlock = one autograd node
_PAIRS_A, _PAIRS_B = (2, 4, 6, 8), (3, 5, 7, 9)


class _TrainBlock:
    """Training view of an IFBlock: its 12 engine layers in unfused tap form (no residual in the epilogue, phase-form heads) and, per
    layer, the tap-form layer that computes its input gradient.

    The layers are BUILT once (`_build`: the per-tap Python loops of ifnet._pack_* / _conv_layer / _convT_phase_layer, ~100 small
    torch launches per layer) and REFRESHED whenever a parameter changes — every training step — from `_Source` records (which
    parameter tensor, in which orientation and tap order, lands where in a layer's tap form): ONE ofsv_conv_refresh_tapform launch
    for the 24 layers of the block, then the library re-packs the bf16 operand blocks (one launch per layer).  On CPU tensors (the
    wiring tests) the same records are applied with torch ops.  tests check refresh == rebuild exactly for both."""

    class _Source:
        """w_simt[:, ci0:ci0+a, co0:co0+b] = W[kidx] with W = param viewed [K][a][b] (perm = the permute of (A, B, K) that gets there);
        kidx < 0 marks a tap the kernel does not have (zero matrix)."""

        def __init__(self, param, perm, kidx, ci0=0, co0=0):
            self.param, self.perm, self.ci0, self.co0 = param, perm, ci0, co0
            self.raw_kidx = list(kidx)
            k = torch.tensor(kidx, device=param.device)
            self.valid = None if bool((k >= 0).all()) else (k >= 0).float().view(-1, 1, 1)
            self.kidx = k.clamp(min=0)

        def apply(self, w_simt):
            p = self.param.detach()
            wk = p.reshape(p.shape[0], p.shape[1], -1).permute(*self.perm).index_select(0, self.kidx)
            if self.valid is not None:
                wk = wk * self.valid
            w_simt[:, self.ci0:self.ci0 + wk.shape[1], self.co0:self.co0 + wk.shape[2]] = wk

    def __init__(self, blk: IFBlock):
        self.blk = blk
        self.key = None
        self.names = [k for k, _ in blk.named_parameters()]
        self.fwd = None

    def _build(self):
        """Structure + initial weights through the reference (slow) packing functions; returns (fwd, dgrad) layer lists."""
        from .ifnet import _pack_conv, _pack_convT
        blk = self.blk
        nd, c, nf = blk.nd, blk.c, 2 * blk.nd
        k0 = blk.conv0[0][0].k
        fwd = [_pack_conv(blk.conv0[0][0], blk.conv0[0][1]), _pack_conv(blk.conv0[1][0], blk.conv0[1][1])]
        for i in range(4):
            cb = getattr(blk, f"convblock{i}")
            fwd += [_pack_conv(cb[0][0], cb[0][1]), _pack_conv(cb[1][0], cb[1][1])]
        wm = torch.cat([blk.conv1[0].weight, blk.conv2[0].weight], 1).detach().float()   # merged ConvT [c][c][4..]
        fwd.append(_pack_convT(nd, wm, torch.cat([blk.conv1[0].bias, blk.conv2[0].bias]).detach().float(),
                               torch.cat([blk.conv1[1].weight, blk.conv2[1].weight]).detach().float(), c, False))
        wh = torch.zeros((c, nf + 1) + (4,) * nd, device=wm.device)                      # block-diagonal heads
        wh[: c // 2, :nf] = blk.conv1[2].weight.detach().float()
        wh[c // 2:, nf:] = blk.conv2[2].weight.detach().float()
        fwd.append(_pack_convT(nd, wh, torch.cat([blk.conv1[2].bias, blk.conv2[2].bias]).detach().float(), None, 8, True))
        dg = [_convT_phase_layer(nd, blk.conv0[0][0].weight.detach().float(), k0, 16),   # conv0.0: c/2 -> block input (16 stored)
              _convT_phase_layer(nd, blk.conv0[1][0].weight.detach().float(), k0, _rup(c // 2, 16))]   # conv0.1: c -> c/2
        for i in range(4):
            cb = getattr(blk, f"convblock{i}")
            for j in range(2):
                dg.append(_conv_layer(nd, cb[j][0].weight.detach().float(), 3, 1, 1, _rup(c, 16), mirror=True))
        dg.append(_conv_layer(nd, wm, 4, 2, 1, _rup(c, 16)))                             # as Conv weight [Cout = c in][Cin = c merged out]
        wh16 = torch.zeros((c, 16) + (4,) * nd, device=wm.device)                        # 16 stored gradient channels
        wh16[:, :nf + 1] = wh
        dg.append(_conv_layer(nd, wh16, 4, 2, 1, _rup(c, 16)))
        return fwd, dg

    def _sources(self):
        """Per layer: ([weight sources], [(bias param, offset)], [(prelu param, offset)]) for fwd, [weight sources] for dgrad."""
        blk = self.blk
        nd, c, nf, h = blk.nd, blk.c, 2 * blk.nd, blk.c // 2
        S = _TrainBlock._Source
        k0 = blk.conv0[0][0].k
        ident = lambda k: list(range(k ** nd))                                           # noqa: E731
        ct = _convT_kflat(nd)
        # Conv(k0, 2, 1) input gradient: phase-form taps of ConvTranspose(k0, 2, 1), None -> -1
        ctk = []
        for par in itertools.product((0, 1), repeat=nd):
            for choice in itertools.product((0, 1), repeat=nd):
                ks = [_CT_TAPS[k0][par[a]][choice[a]][0] for a in range(nd)]
                f = -1
                if all(v is not None for v in ks):
                    f = 0
                    for v in ks:
                        f = f * k0 + v
                ctk.append(f)
        AB, BA = (2, 0, 1), (2, 1, 0)                                                    # param (A,B,K) -> [K][A][B] / [K][B][A]
        fw, dg = [], []
        for li in (0, 1):
            m, pr = blk.conv0[li][0], blk.conv0[li][1]
            fw.append(([S(m.weight, BA, ident(k0))], [(m.bias, 0)], [(pr.weight, 0)]))
            dg.append([S(m.weight, AB, ctk)])
        for i in range(4):
            cb = getattr(blk, f"convblock{i}")
            for j in range(2):
                fw.append(([S(cb[j][0].weight, BA, ident(3))], [(cb[j][0].bias, 0)], [(cb[j][1].weight, 0)]))
                dg.append([S(cb[j][0].weight, AB, ident(3))])
        fw.append(([S(blk.conv1[0].weight, AB, ct), S(blk.conv2[0].weight, AB, ct, co0=h)],
                   [(blk.conv1[0].bias, 0), (blk.conv2[0].bias, h)], [(blk.conv1[1].weight, 0), (blk.conv2[1].weight, h)]))
        dg.append([S(blk.conv1[0].weight, BA, ident(4)), S(blk.conv2[0].weight, BA, ident(4), ci0=h)])
        fw.append(([S(blk.conv1[2].weight, AB, ct), S(blk.conv2[2].weight, AB, ct, ci0=h, co0=nf)],
                   [(blk.conv1[2].bias, 0), (blk.conv2[2].bias, nf)], []))
        dg.append([S(blk.conv1[2].weight, BA, ident(4)), S(blk.conv2[2].weight, BA, ident(4), ci0=nf, co0=h)])
        return fw, dg

    def _record_table(self):
        """Device table of ofsv_refresh_rec: every weight source and bias / slope vector of the 24 layers, for ONE refresh launch."""
        recs = []

        def weight(lay, sct):
            r = _C.RefreshRec()
            p = sct.param
            r.src, r.dst, r.kind = p.data_ptr(), lay.w_simt.data_ptr(), 0
            r.A, r.B, r.K = p.shape[0], p.shape[1], p[0, 0].numel()
            r.swap = 0 if tuple(sct.perm) == (2, 0, 1) else 1
            r.T, r.Cin_s, r.Cout_w = lay.w_simt.shape
            r.ci0, r.co0 = sct.ci0, sct.co0
            for i, k in enumerate(sct.raw_kidx):
                r.kidx[i] = k
            recs.append(r)

        def vector(dst, off, p):
            r = _C.RefreshRec()
            r.src, r.dst, r.kind, r.n = p.data_ptr(), dst.data_ptr() + 4 * off, 1, p.numel()
            recs.append(r)

        for lay, (ws, bs, ps) in zip(self.fwd, self.src_fwd):
            for sct in ws:
                weight(lay, sct)
            for b, off in bs:
                vector(lay.bias, off, b)
            for pw, off in ps:
                vector(lay.prelu, off, pw)
        for lay, ws in zip(self.dgrad, self.src_dgrad):
            for sct in ws:
                weight(lay, sct)
        raw = b"".join(bytes(r) for r in recs)
        dev = self.fwd[0].w_simt.device
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev), len(recs)

    def _repack(self):
        """bf16 operand blocks of every (layer, layout) that exists, re-packed IN PLACE by one launch (ofsv_conv_pack_weights_batched).
        The table follows the set of packed forms (the first step creates them lazily, layer by layer, and decides halo vs per-tap
        there); building it is a host-to-device copy, so the graph path refreshes once eagerly before it captures."""
        layers = self.fwd + self.dgrad
        keys = tuple(tuple(sorted(lay._packed)) for lay in layers)          # a layer that has not run yet has no packed form: the
        if getattr(self, "_pack_table", None) is None or self._pack_keys != keys:    # engines pack it lazily, from the current tap form
            recs = []
            L = _C.lib()
            for lay in layers:
                d = lay._structure_desc()
                for layout, w_out in sorted(lay._packed.items()):
                    r = _C.PackRec()
                    _C.check(L.ofsv_conv_pack_record(ctypes.byref(d), int(layout), _p(lay.w_simt), _p(w_out), ctypes.byref(r)))
                    recs.append(r)
            raw = b"".join(bytes(r) for r in recs) or b"\0" * 8
            self._pack_table = (torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(layers[0].w_simt.device), len(recs))
            self._pack_keys = keys
        tab, n = self._pack_table
        with ops._on(tab.device), ops._span("conv_pack_batched"):
            _C.check(_C.lib().ofsv_conv_pack_weights_batched(_p(tab), n, _stream()))

    def refresh(self):
        key = self.blk._key()
        if key == self.key:
            return
        if self.fwd is not None and self.fwd[0].w_simt.device != next(self.blk.parameters()).device:
            self.fwd = None                                  # the module was moved to another device: build there from scratch
        with torch.no_grad():
            if self.fwd is None:
                self.fwd, self.dgrad = self._build()
                self.src_fwd, self.src_dgrad = self._sources()
                self.kinv = torch.tensor(_convT_kflat(self.blk.nd), device=self.fwd[0].w_simt.device).argsort()
                self._table, self._table_ptrs, self._pack_table = None, None, None
                if self.fwd[0].w_simt.is_cuda:          # built here, outside any CUDA-graph capture (it is a host-to-device copy)
                    self._table, self._table_ptrs = self._record_table(), tuple(p.data_ptr() for p in self.blk.parameters())
            elif self.fwd[0].w_simt.is_cuda:
                ptrs = tuple(p.data_ptr() for p in self.blk.parameters())
                if self._table is None or self._table_ptrs != ptrs:         # parameters moved (load_state_dict keeps them, .to() may not)
                    self._table, self._table_ptrs = self._record_table(), ptrs
                tab, n = self._table
                with ops._on(tab.device), ops._span("conv_refresh"):
                    _C.check(_C.lib().ofsv_conv_refresh_tapform(_p(tab), n, _stream()))
                self._repack()
            else:                                                           # tests/test_train_host.py: the same refresh with torch ops
                for lay, (ws, bs, ps) in zip(self.fwd, self.src_fwd):
                    for sct in ws:
                        sct.apply(lay.w_simt)
                    for b, off in bs:
                        lay.bias[off:off + b.numel()] = b.detach()
                    for pw, off in ps:
                        lay.prelu[off:off + pw.numel()] = pw.detach()
                    lay._packed.clear()
                for lay, ws in zip(self.dgrad, self.src_dgrad):
                    for sct in ws:
                        sct.apply(lay.w_simt)
                    lay._packed.clear()
        self.key = key

    # -- tap-form weight gradients back to the reference's parameter tensors
    def conv_weight_grad(self, dw, m):
        """dw [k^nd][Cin_s][Cout_w] of ifnet._pack_conv -> grad of nn.Conv weight [Cout][Cin][k..]."""
        return dw[:, :m.cin, :m.cout].permute(2, 1, 0).reshape(m.weight.shape)

    def convT_weight_grad(self, dw, cin, cout):
        """dw [4^nd (parity, choice)][Cin_s][Cout_w] of ifnet._pack_convT -> [cin][cout][4..]."""
        nd = self.blk.nd
        return dw[self.kinv][:, :cin, :cout].permute(1, 2, 0).reshape((cin, cout) + (4,) * nd)


class _BlockFn(torch.autograd.Function):
    """head = IFBlock convs(x): x fp32 (N, Cin, *sp) at block resolution -> head fp32 (N, 2nd+1, *sp) (flow delta ‖ mask delta before
    the up-resize).  Flow-3D/model/IFNet.py:91-116 / Flow-2D/model/IFNet.py:95-113 (conv0, convblock0-3 with skips, conv1, conv2)."""

    @staticmethod
    def forward(ctx, x, tb, cl_io, *params):
        """cl_io = False: x fp32 (N, Cin, *sp), returns the head fp32 (N, 2nd+1, *sp).  cl_io = True: x is the packed block input itself
        (bf16 channels-last [N][D][H][W][16], ofsv_pack_block_input) and the head comes back as the engine wrote it (fp32
        channels-last [N][D][H][W][8]) — no layout copies on either side."""
        _require_cuda(x, "IFBlock training forward")
        blk = tb.blk
        nd = blk.nd
        tb.refresh()
        n = x.shape[0]
        sp3 = tuple(x.shape[1:4]) if cl_io else (((1,) + tuple(x.shape[2:])) if nd == 2 else tuple(x.shape[2:]))
        with ops._on(x.device):
            a = x.detach() if cl_io else _to_cl16(x.detach(), nd)
            xs, ys, descs = [], [], []
            cur, cur_sp, skip = a, sp3, None
            for li, lay in enumerate(tb.fwd):
                d, osp = lay.desc(n, cur_sp, _C.BF16, has_residual=False)
                y = torch.empty(lay.out_shape(n, osp), device=x.device, dtype=torch.float32 if lay.out_f32 else _ACT_DTYPE)
                if nd == 2:
                    y = y.view(n, 1, *y.shape[1:])
                _run_layer(lay, d, cur, None, y)
                xs.append(cur); ys.append(y); descs.append(d)
                if li in _PAIRS_A:
                    skip = cur
                cur = (y + skip) if li in _PAIRS_B else y
                cur_sp = osp
            head = ys[-1] if cl_io else _from_cl(ys[-1], 2 * nd + 1, nd)
        ctx.tb, ctx.xs, ctx.ys, ctx.descs, ctx.nd, ctx.n = tb, xs, ys, descs, nd, n
        ctx.cin = 16 if cl_io else x.shape[1]
        ctx.need_x = x.requires_grad
        ctx.in_sp, ctx.cl_io = sp3, cl_io
        return head

    @staticmethod
    def backward(ctx, g_head):
        tb, xs, ys, descs, nd, n = ctx.tb, ctx.xs, ctx.ys, ctx.descs, ctx.nd, ctx.n
        blk = tb.blk
        c, nf = blk.c, 2 * nd
        grads = {}
        dev = g_head.device

        def dgrad(li, g, in_sp, res=None):
            lay = tb.dgrad[li]
            d, osp = lay.desc(n, in_sp, _C.BF16, has_residual=res is not None)
            y = torch.empty(lay.out_shape(n, osp), device=dev, dtype=_ACT_DTYPE)
            if nd == 2:
                y = y.view(n, 1, *y.shape[1:])
            _run_layer(lay, d, g, res, y)
            return y

        def sp_of(t):
            return tuple(t.shape[1:4])

        with ops._on(dev):
            g_head = g_head.contiguous()
            # heads (merged block-diagonal ConvT c -> 2nd+1, no activation): bias gradient is a plain sum
            if ctx.cl_io:                                     # fp32 channels-last [N][D][H][W][8] -> bf16 [..][16]
                db = g_head.reshape(-1, 8).sum(0)
                g = torch.zeros(g_head.shape[:-1] + (16,), device=dev, dtype=_ACT_DTYPE)
                g[..., :8] = g_head
            else:
                db = g_head.sum((0,) + tuple(range(2, 2 + nd)))
                g = _to_cl16(g_head, nd)
            grads["conv1.2.bias"], grads["conv2.2.bias"] = db[:nf], db[nf:nf + 1]
            dw = tb.convT_weight_grad(conv_wgrad(descs[11], xs[11], g), c, nf + 1)
            grads["conv1.2.weight"], grads["conv2.2.weight"] = dw[: c // 2, :nf], dw[c // 2:, nf:nf + 1]
            g = dgrad(11, g, sp_of(g))
            # merged conv1.0 ‖ conv2.0 + PReLU
            gp, db, ds = prelu_bias_bwd(g, ys[10], tb.fwd[10].prelu)
            h = c // 2
            grads["conv1.0.bias"], grads["conv2.0.bias"] = db[:h], db[h:c]
            grads["conv1.1.weight"], grads["conv2.1.weight"] = ds[:h], ds[h:c]
            dw = tb.convT_weight_grad(conv_wgrad(descs[10], xs[10], gp), c, c)
            grads["conv1.0.weight"], grads["conv2.0.weight"] = dw[:, :h], dw[:, h:]
            g = dgrad(10, gp, sp_of(gp))
            # residual pairs, last to first:  out = PReLU(conv_b(PReLU(conv_a(x)))) + x
            for i in (3, 2, 1, 0):
                la, lb = 2 + 2 * i, 3 + 2 * i
                cb = getattr(blk, f"convblock{i}")
                g_out = g
                gp, db, ds = prelu_bias_bwd(g_out, ys[lb], tb.fwd[lb].prelu)
                grads[f"convblock{i}.1.0.bias"], grads[f"convblock{i}.1.1.weight"] = db[:c], ds[:c]
                grads[f"convblock{i}.1.0.weight"] = tb.conv_weight_grad(conv_wgrad(descs[lb], xs[lb], gp), cb[1][0])
                g = dgrad(lb, gp, sp_of(gp))
                gp, db, ds = prelu_bias_bwd(g, ys[la], tb.fwd[la].prelu)
                grads[f"convblock{i}.0.0.bias"], grads[f"convblock{i}.0.1.weight"] = db[:c], ds[:c]
                grads[f"convblock{i}.0.0.weight"] = tb.conv_weight_grad(conv_wgrad(descs[la], xs[la], gp), cb[0][0])
                g = dgrad(la, gp, sp_of(gp), res=g_out)
            # conv0.1, conv0.0
            for li in (1, 0):
                m = blk.conv0[li][0]
                gp, db, ds = prelu_bias_bwd(g, ys[li], tb.fwd[li].prelu)
                grads[f"conv0.{li}.0.bias"], grads[f"conv0.{li}.1.weight"] = db[:m.cout], ds[:m.cout]
                grads[f"conv0.{li}.0.weight"] = tb.conv_weight_grad(conv_wgrad(descs[li], xs[li], gp), m)
                if li == 1 or ctx.need_x:
                    g = dgrad(li, gp, sp_of(gp))
            gx = (g if ctx.cl_io else _from_cl(g, ctx.cin, nd)) if ctx.need_x else None
        return (gx, None, None) + tuple(grads[k].contiguous() for k in tb.names)


def _warp_fn(nd):
    return ops.warp2d if nd == 2 else ops.warp3d


def block_train(tb: _TrainBlock, x, flow, scale):
    """IFBlock.forward (Flow-3D/model/IFNet.py:80-119, Flow-2D/model/IFNet.py:84-116) with the conv stack as one libofsv node."""
    nd = tb.blk.nd
    mode = "bilinear" if nd == 2 else "trilinear"
    if scale != 1:
        x = F.interpolate(x, scale_factor=1. / scale, mode=mode, align_corners=False)
    if flow is not None:
        flow = F.interpolate(flow, scale_factor=1. / scale, mode=mode, align_corners=False) * 1. / scale
        x = torch.cat((x, flow), 1)
    head = _BlockFn.apply(x, tb, False, *tb.blk.parameters())
    nf = 2 * nd
    flow_d = F.interpolate(head[:, :nf], scale_factor=scale, mode=mode, align_corners=False, recompute_scale_factor=False) * scale
    mask_d = F.interpolate(head[:, nf:nf + 1], scale_factor=scale, mode=mode, align_corners=False, recompute_scale_factor=False)
    return flow_d, mask_d


def _pack_input(img0, img1, w0, w1, mask, flow, scale):
    """ofsv_pack_block_input in the plain layout, always 5-D: [N][D][H][W][16] with D = 1 for frames."""
    xin = ops.pack_block_input(img0, img1, w0, w1, mask, flow, scale, _C.BF16)
    return xin.unsqueeze(1) if img0.dim() == 4 else xin


class _PackInputFn(torch.autograd.Function):
    """xin = ofsv_pack_block_input(img0, img1, warped0, warped1, mask, flow; scale): resize 1/s of the concatenation with the flow / s
    (IFNet.py:84-93 / :82-90 + the cat of :174 / :166), bf16 channels-last [N][D][H][W][16]; backward: ofsv_pack_block_input_bwd."""

    @staticmethod
    def forward(ctx, img0, img1, w0, w1, mask, flow, scale):
        ctx.scale, ctx.nd, ctx.sp = scale, img0.dim() - 2, tuple(img0.shape[2:])
        return _pack_input(img0, img1, w0.contiguous(), w1.contiguous(), mask.contiguous(), flow.contiguous(), scale)

    @staticmethod
    def backward(ctx, gx):
        g0, g1, gm, gf = ops.pack_block_input_bwd(gx.contiguous(), ctx.nd, ctx.sp, ctx.scale)
        return None, None, g0, g1, gm, gf, None


class _PackCatFn(torch.autograd.Function):
    """The teacher block's input torch.cat((img0, img1, warped0, warped1, mask, gt, flow), 1) at scale 1 (IFNet.py:215 / :213; the
    resizes of IFBlock.forward are identities there) as bf16 channels-last [N][D][H][W][16] in one launch (ofsv_pack_nhwc_bf16);
    backward: ofsv_unpack_nhwc_f32 of the input gradient, split back into the sources."""

    @staticmethod
    def forward(ctx, *srcs):
        ctx.chans = [t.shape[1] for t in srcs]
        ctx.nd = srcs[0].dim() - 2
        return ops.pack_nhwc([t.contiguous() for t in srcs], 16)

    @staticmethod
    def backward(ctx, gx):
        g = ops.unpack_nhwc(gx.contiguous(), sum(ctx.chans), ctx.nd)
        out, off = [], 0
        for i, c in enumerate(ctx.chans):
            out.append(g[:, off:off + c] if ctx.needs_input_grad[i] else None)
            off += c
        return tuple(out)


class _HeadUpFn(torch.autograd.Function):
    """(flow, mask) = (flow_prev + s * up_s(head[:2nd]), mask_prev + up_s(head[2nd])) — IFNet.py:115-119,177-178 / :118-119,169-170 — on
    the fp32 channels-last head (ofsv_head_upsample_add); backward: ofsv_head_upsample_add_bwd (+ identity to the previous state)."""

    @staticmethod
    def forward(ctx, head, flow_prev, mask_prev, scale, nd, sp):
        ctx.scale, ctx.nd, ctx.has_prev = scale, nd, flow_prev is not None
        fp = flow_prev.contiguous() if flow_prev is not None else None
        mp = mask_prev.contiguous() if mask_prev is not None else None
        flow, mask = ops.head_upsample_add(head, fp, mp, nd, head.shape[0], sp, scale)
        return flow, mask

    @staticmethod
    def backward(ctx, gflow, gmask):
        gflow, gmask = gflow.contiguous(), gmask.contiguous()
        ghead = ops.head_upsample_add_bwd(gflow, gmask, ctx.nd, ctx.scale)
        return ghead, (gflow if ctx.has_prev else None), (gmask if ctx.has_prev else None), None, None, None


def ifnet_forward_train(net: IFNet, x, scale=(4, 2, 1)):
    """`IFNet.forward` with gt as the third channel (Flow-3D/model/IFNet.py:133-280, Flow-2D/model/IFNet.py:144-276): three
    student blocks, the teacher block on (…, gt), the distillation mask and loss.  Returns the reference's tuple
    (flow_list, mask_list[2] | mask_list, merged, flow_teacher, merged_teacher, loss_distill)."""
    nd = net.nd
    if net.precision != "bf16":
        raise NotImplementedError("the training step runs on the bf16 tensor-core engine only")
    if x.dim() != nd + 2 or x.shape[1] != 3:
        raise ValueError(f"training forward: expected cat(img0, img1, gt) of shape (N,3,{'D,' if nd == 3 else ''}H,W), got {tuple(x.shape)}")
    if any(s % 16 for s in x.shape[2:]):
        raise NotImplementedError("spatial dims must be multiples of 16 (the reference's shape-repair slicing is not reproduced)")
    x = _require_cuda(x, "x").float()
    warp = _warp_fn(nd)
    tbs = getattr(net, "_train_blocks", None)
    if tbs is None:
        tbs = net._train_blocks = [_TrainBlock(b) for b in (net.block0, net.block1, net.block2, net.block_tea)]
    img0, img1, gt = x[:, :1].contiguous(), x[:, 1:2].contiguous(), x[:, 2:3].contiguous()
    flow_list, mask_list, warped = [], [], []
    w0, w1, flow, mask = img0, img1, None, None
    fused = x.is_cuda          # student blocks: input packing and head up-sampling + accumulate as libofsv nodes (no torch resizes)
    sp = tuple(x.shape[2:])
    for i in range(3):
        if fused:
            s = int(scale[i])
            if flow is None:
                xin = _pack_input(img0, img1, None, None, None, None, s)
            else:
                xin = _PackInputFn.apply(img0, img1, w0, w1, mask, flow, s)
            head = _BlockFn.apply(xin, tbs[i], True, *tbs[i].blk.parameters())
            flow, mask = _HeadUpFn.apply(head, flow, mask, s, nd, sp)
        elif flow is None:
            flow, mask = block_train(tbs[i], torch.cat((img0, img1), 1), None, scale[i])
        else:
            fd, md = block_train(tbs[i], torch.cat((img0, img1, w0, w1, mask), 1), flow, scale[i])
            flow, mask = flow + fd, mask + md
        mask_list.append(torch.sigmoid(mask))
        flow_list.append(flow)
        w0 = warp(img0, flow[:, :nd].contiguous())
        w1 = warp(img1, flow[:, nd:2 * nd].contiguous())
        warped.append((w0, w1))
    if fused:
        xin = _PackCatFn.apply(img0, img1, w0, w1, mask, gt, flow)
        head = _BlockFn.apply(xin, tbs[3], True, *tbs[3].blk.parameters())
        flow_teacher, mask_tea_logit = _HeadUpFn.apply(head, flow, mask, 1, nd, sp)
        fd = md = None
    else:
        fd, md = block_train(tbs[3], torch.cat((img0, img1, w0, w1, mask, gt), 1), flow, 1)
        flow_teacher, mask_tea_logit = flow + fd, mask + md
    w0t = warp(img0, flow_teacher[:, :nd].contiguous())
    w1t = warp(img1, flow_teacher[:, nd:2 * nd].contiguous())
    mask_teacher = torch.sigmoid(mask_tea_logit)
    merged_teacher = w0t * mask_teacher + w1t * (1 - mask_teacher)
    merged, loss_distill = [], 0
    for i in range(3):
        merged.append(warped[i][0] * mask_list[i] + warped[i][1] * (1 - mask_list[i]))
        loss_mask = ((merged[i] - gt).abs().mean(1, True) > (merged_teacher - gt).abs().mean(1, True) + 0.01).float().detach()
        loss_distill = loss_distill + (((flow_teacher.detach() - flow_list[i]) ** 2).mean(1, True) ** 0.5 * loss_mask).mean()
    return flow_list, (mask_list if nd == 2 else mask_list[2]), merged, flow_teacher, merged_teacher, loss_distill


# ------------------------------------------------------------------------------------------------ 2-D losses (Flow-2D/model)
_CONST = {}


def _const(key, device, make):
    """Small constant tensors are built once per device (a host-to-device copy is not allowed inside CUDA-graph capture)."""
    k = (key, str(device))
    t = _CONST.get(k)
    if t is None:
        t = _CONST[k] = make().to(device)
    return t


def _gauss_kernel(channels, device):
    """Flow-2D/model/laplacian.py:10-19."""
    def make():
        k = torch.tensor([1., 4., 6., 4., 1.])
        return (torch.outer(k, k) / 256.).repeat(channels, 1, 1, 1)
    return _const(("gauss", channels), device, make)


def _conv_gauss(img, kernel):
    return F.conv2d(F.pad(img, (2, 2, 2, 2), mode="reflect"), kernel, groups=img.shape[1])


def _lap_upsample(x):
    """laplacian.py:24-31: zero-interleave to 2x, then 4 * Gaussian."""
    n, c, h, w = x.shape
    up = torch.zeros(n, c, 2 * h, 2 * w, device=x.device, dtype=x.dtype)
    up[:, :, ::2, ::2] = x
    return _conv_gauss(up, 4 * _gauss_kernel(c, x.device))


def lap_loss(inp, target, max_levels=5):
    """LapLoss(max_levels=5, channels=1) — laplacian.py:38-75."""
    kernel = _gauss_kernel(inp.shape[1], inp.device)

    def pyramid(img):
        cur, pyr = img, []
        for _ in range(max_levels):
            down = _conv_gauss(cur, kernel)[:, :, ::2, ::2]
            up = _lap_upsample(down)
            h, w = min(cur.shape[2], up.shape[2]), min(cur.shape[3], up.shape[3])
            pyr.append(cur[:, :, :h, :w] - up[:, :, :h, :w])
            cur = down
        return pyr

    return sum(F.l1_loss(a, b) for a, b in zip(pyramid(inp), pyramid(target)))


def _charbonnier(x, alpha=0.25, epsilon=1.e-9):
    return torch.pow(torch.pow(x, 2) + epsilon ** 2, alpha)


def _photometric_term(flow, merged, frame):
    """Flow-2D/model/RIFE.py:245-282: backwrd_warp (grid_sample with zeros padding / align_corners=False on a (2/w, 2/h) grid) of
    `merged` by `flow`, Charbonnier distance to `frame`, summed over pixels / 3 / batch."""
    b, _, h, w = flow.shape
    def base():
        yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        return torch.stack((xx, yy), -1).float().unsqueeze(0)
    grid = flow.permute(0, 2, 3, 1) + _const(("photo_grid", h, w), flow.device, base)
    grid = grid * _const(("photo_factor", h, w), flow.device, lambda: torch.tensor([2 / w, 2 / h])) - 1
    # loss glue with weight 1e-5, not on the inference path: ATen's sampler (zeros padding, align_corners=False) is used as is
    warped = F.grid_sample(merged, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    p = _charbonnier(warped - frame)
    return torch.sum(torch.sum(p, dim=1) / 3) / b


# ------------------------------------------------------------------------------------------------ Model.update
class Trainer:
    """Optimizer state + the step itself; owned by `rife.Model{2,3}D` (created on the first `update`)."""

    def __init__(self, net: IFNet, local_rank=-1):
        from .optim import FusedAdamW, GradientBucket
        self.net = net
        self.params = [p for p in net.parameters()]
        self.bucket = GradientBucket(self.params)
        self.optimG = FusedAdamW(self.params, lr=1e-6, weight_decay=1e-3, bucket=self.bucket)   # RIFE.py:29 / :26
        self.distributed = local_rank != -1
        self.allreduce_events = None      # bench.py: a list collects (start, stop) CUDA events around the gradient all-reduce
        self._slopes = [p for k, p in net.named_parameters() if p.dim() == 1 and k.endswith(".1.weight")]
        self.graphs = None                # {(shapes): (CUDAGraph, static imgs, static gt, static outputs)} when graph mode is on

    def check_slopes(self):
        """ofsv_prelu_bias_bwd_bf16 recovers the pre-activation sign from the layer output, which needs PReLU slopes > 0 (the
        reference initialises 0.25).  The minimum is computed on the device after every step and read at the start of the NEXT
        one, when it is long finished (no pipeline stall)."""
        m = getattr(self, "_min_slope", None)
        if m is not None and not (m.item() > 0):
            raise RuntimeError("a PReLU slope became <= 0: the fused bias/PReLU backward of the bf16 training path needs positive slopes")

    def forward_backward(self, imgs, gt, device_guard=False):
        """zero_grad + forward + losses + backward (everything of the step before the gradient exchange)."""
        self.bucket.zero()                                   # optimG.zero_grad()
        with torch.enable_grad():
            loss_G, info, merged2 = forward_losses(self.net, imgs, gt, device_guard)
            loss_G.backward()
        return merged2.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in info.items() if k != "_flow2"}

    def forward_backward_graphed(self, imgs, gt):
        """The same work replayed from a CUDA graph captured per input shape: a step is ~400 libofsv launches plus ~300 small torch
        launches, and at the reference's training sizes (64^3, 128^2 ...) the GPU finishes them faster than Python can enqueue
        them.  The graph contains the per-step weight refresh, so it stays valid while the optimizer updates the parameters in
        place; inputs are copied into the graph's buffers and the returned tensors are the graph's outputs, overwritten by the next
        step (the small loss scalars are cloned).  The gradient all-reduce and the optimizer stay outside the graph: the learning
        rate changes every step (RIFE.py:86-87)."""
        key = (tuple(imgs.shape), tuple(gt.shape), imgs.device.index)
        ent = self.graphs.get(key)
        if ent is None:
            s_imgs, s_gt = imgs.detach().clone().contiguous(), gt.detach().clone().contiguous()
            cur = torch.cuda.current_stream(imgs.device)
            side = torch.cuda.Stream(device=imgs.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                     # warm-up off the capture: workspaces, kernel attributes, engine choice
                for _ in range(2):
                    self.forward_backward(s_imgs, s_gt, device_guard=True)
            cur.wait_stream(side)
            for tb in self.net._train_blocks:                 # tables complete (built outside the capture) ...
                tb.key = None
                tb.refresh()
            for tb in self.net._train_blocks:                 # ... and the capture must contain the weight refresh
                tb.key = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward_backward(s_imgs, s_gt, device_guard=True)
            ent = self.graphs[key] = (g, s_imgs, s_gt, out)
        g, s_imgs, s_gt, (merged2, info) = ent
        s_imgs.copy_(imgs)
        s_gt.copy_(gt)
        g.replay()
        return merged2, {k: (v.clone() if torch.is_tensor(v) and v.dim() == 0 else v) for k, v in info.items()}

    def exchange_and_step(self):
        from .optim import allreduce_gradients
        ev = self.allreduce_events
        if ev is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        scale = allreduce_gradients(self.bucket) if self.distributed else 1.0
        if ev is not None:
            b.record()
            ev.append((a, b))
        self.optimG.step(grad_scale=scale)
        with torch.no_grad():
            self._min_slope = torch.cat([p.view(-1) for p in self._slopes]).min()


def forward_losses(net, imgs, gt, device_guard=False):
    """Forward with the teacher block and the losses of `update`.  device_guard: replace the 2-D path's host-side
    `isnan(loss_distill) or loss_distill > 10` test (RIFE.py:295) by the same selection on the device (needed under CUDA-graph
    capture; when it triggers, the reference drops the term from the graph while this form multiplies its gradient by zero)."""
    nd = net.nd
    img0, img1 = imgs[:, :1], imgs[:, 1:2]
    flow, mask, merged, flow_teacher, merged_teacher, loss_distill = ifnet_forward_train(net, torch.cat((imgs, gt), 1), (4, 2, 1))
    if nd == 3:
        loss_l1 = F.l1_loss(merged[2], gt)
        loss_tea = F.l1_loss(merged_teacher, gt)
        loss_G = loss_l1 * 1 + loss_tea * 1 + loss_distill * 0.1
        info = {"loss_l1": loss_l1, "loss_tea": loss_tea, "loss_distill": loss_distill, "loss_G": loss_G}
    else:
        mask = mask[2]
        loss_l1 = lap_loss(merged[2], gt).mean()
        loss_tea = lap_loss(merged_teacher, gt).mean()
        with torch.no_grad():
            l1_reg = sum(torch.norm(p, 1) for k, p in net.state_dict().items() if "block2" in k or "block_tea" in k)
        loss_photo = (_photometric_term(flow[2][:, 2:4], merged[2], img0) + _photometric_term(flow[2][:, :2], merged[2], img1)) / 2
        if device_guard:
            loss_distill = torch.where(torch.isnan(loss_distill) | (loss_distill > 10.), torch.zeros_like(loss_distill), loss_distill)
        else:
            ld = float(loss_distill.detach())                              # host read, as in the reference (RIFE.py:295)
            if math.isnan(ld) or ld > 10.:
                loss_distill = torch.zeros((), device=imgs.device)
        loss_G = loss_l1 * 1 + loss_tea * 1 + loss_distill * 0.01 + l1_reg * 1e-6 + loss_photo * 1e-5
        info = {"loss_l1": loss_l1 * 1, "loss_tea": loss_tea * 1, "loss_distill": loss_distill * 0.01, "l1_reg": l1_reg * 1e-6,
                "loss_photo": loss_photo * 1e-5, "loss_flow": torch.zeros(()) * 0, "loss_G": loss_G}
    info.update({"merged_tea": merged_teacher, "mask": mask, "mask_tea": mask, "flow": flow[2] if nd == 3 else flow[2][:, :2],
                 "flow_tea": flow_teacher, "_flow2": flow[2]})
    return loss_G, info, merged[2]


def update(model, imgs, gt, learning_rate=0, mul=1, training=True, flow_gt=None, dataset=None):
    """`Model.update` of both packages.  3-D (Flow-3D/model/RIFE.py:81-275): loss_G = L1(merged[2], gt) + L1(merged_teacher, gt)
    + 0.1 * loss_distill.  2-D (Flow-2D/model/RIFE.py:80-336), 1-channel datasets (`droplet2d`, `vimeo2d`): LapLoss student and
    teacher, 0.01 * distillation (zeroed when NaN or > 10), 1e-6 * |w|_1 of block2 / block_tea (a constant for autograd: the
    reference reads it through state_dict()), 1e-5 * photometric loss.  Returns (merged[2], info dict with the reference's keys).
    `model.enable_training_graph()` replays forward + backward from a CUDA graph (see Trainer.forward_backward_graphed)."""
    net = model.flownet
    nd = net.nd
    if nd == 2 and dataset not in (None, "droplet2d", "vimeo2d"):
        raise NotImplementedError("2-D update: only the 1-channel dataset branch (droplet2d / vimeo2d) is provided; the "
                                  "data+flow-channel datasets of RIFE.py:86-103 are outside the hot path")
    for t, name in ((imgs, "imgs"), (gt, "gt")):
        if not torch.is_tensor(t) or not t.is_cuda:
            raise TypeError(f"{name}: expected a CUDA tensor (no CPU path)")
    if net.precision != "bf16":
        raise NotImplementedError("the training step runs on the bf16 tensor-core engine only (Model(precision='bf16'))")
    tr = getattr(model, "_trainer", None)
    if tr is None:
        tr = model._trainer = Trainer(net, model.local_rank)
    for g in tr.optimG.param_groups:
        g["lr"] = learning_rate
    tr.check_slopes()
    model.train() if training else model.eval()
    if not training:
        with torch.no_grad():
            _, info, merged2 = forward_losses(net, imgs, gt)
        info["flow_tea"], info["merged_tea"] = info.pop("_flow2"), merged2       # RIFE.py:260-262 / :319-321
        return merged2, info
    want_graph = getattr(model, "_train_graph", False)
    if want_graph and tr.graphs is None:
        tr.graphs = {}
    merged2, info = tr.forward_backward_graphed(imgs, gt) if want_graph else tr.forward_backward(imgs, gt)
    tr.exchange_and_step()
    return merged2, info
