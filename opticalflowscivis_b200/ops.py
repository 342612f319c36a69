"""Tensor-level wrappers over the libofsv C ABI.  PyTorch is plumbing only: it owns device memory and the stream.

Every function takes CUDA tensors, allocates outputs with torch.empty and enqueues the kernel on the current
stream.  CPU tensors raise TypeError — there is deliberately no CPU or ATen fallback.
"""
from __future__ import annotations

import ctypes
import threading

import torch

from . import _C

_FLAVOR = {"mode": _C.REF_CPU}
_LIN = {}


class LaunchTimer:
    """CUDA-event stopwatch per kernel class, recorded on the launching stream (bench.py's roofline numbers).
    Install with `ops.TIMER = LaunchTimer()`; every wrapper below brackets its launch with two events."""

    def __init__(self):
        self.spans = {}
        self.seq = []          # (class, start event, stop event) in launch order

    class _Span:
        def __init__(self, owner, name):
            self.o, self.name = owner, name

        def __enter__(self):
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
            return self

        def __exit__(self, *exc):
            self.b.record()
            self.o.spans.setdefault(self.name, []).append((self.a, self.b))
            self.o.seq.append((self.name, self.a, self.b))
            return False

    def span(self, name):
        return LaunchTimer._Span(self, name)

    def totals(self):
        """{class: (launches, total_ms)} — synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}


class _NoSpan:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


TIMER = None
_NOSPAN = _NoSpan()


def _span(name):
    return TIMER.span(name) if TIMER is not None else _NOSPAN


def set_reference_flavor(flavor: str) -> None:
    """Which build of the reference the fp32 warp arithmetic reproduces bit-for-bit.

    'cpu'  (default): the reference run on CPU — the oracle of this repo (true division, CPU torch.linspace table).
    'cuda': the reference run in CUDA eager (ATen's reciprocal-multiply division, CUDA torch.linspace table).
    The two differ by ~1e-5 max-abs in warped intensities."""
    if flavor not in ("cpu", "cuda"):
        raise ValueError("flavor must be 'cpu' or 'cuda'")
    _FLAVOR["mode"] = _C.REF_CPU if flavor == "cpu" else _C.REF_CUDA


def reference_flavor() -> str:
    return "cpu" if _FLAVOR["mode"] == _C.REF_CPU else "cuda"


def _on(dev):
    """Device guard for a launch: a no-op when `dev` is already current (torch.cuda.device costs ~10 us per launch, more than the
    GPU time of the small layers of a 160x224 frame pair)."""
    if dev.index is None or torch.cuda.current_device() == dev.index:
        return _NOSPAN
    return torch.cuda.device(dev)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor, got {t.device} (this package has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def linspace_table(n: int, device: torch.device) -> torch.Tensor:
    """torch.linspace(-1, 1, n) produced on the device class the selected reference flavour would use
    (its bits differ between ATen's CPU and CUDA kernels), cached on `device`."""
    key = (n, str(device), _FLAVOR["mode"])
    t = _LIN.get(key)
    if t is None:
        src = "cpu" if _FLAVOR["mode"] == _C.REF_CPU else device
        t = torch.linspace(-1.0, 1.0, n, device=src).to(device)
        _LIN[key] = t
    return t


# ------------------------------------------------------------------------------------------------ warp
def _check_warp_shapes(x, f, nd):
    if x.dim() != nd + 2 or f.dim() != nd + 2 or f.shape[1] != nd or f.shape[0] != x.shape[0] or f.shape[2:] != x.shape[2:]:
        raise ValueError(f"warp{nd}d: bad shapes {tuple(x.shape)} / {tuple(f.shape)}")


def _warp_fwd(x, f):
    nd = x.dim() - 2
    dev = x.device
    out = torch.empty_like(x)
    with _on(dev):
        if nd == 2:
            n, c, h, w = x.shape
            _C.check(_C.lib().ofsv_warp2d_f32(_p(x), _p(f), _p(linspace_table(w, dev)), _p(linspace_table(h, dev)),
                                              _p(out), n, c, h, w, _FLAVOR["mode"], _stream()))
        else:
            n, c, d, h, w = x.shape
            with _span("warp3d"):
                _C.check(_C.lib().ofsv_warp3d_f32(_p(x), _p(f), _p(linspace_table(h, dev)), _p(linspace_table(d, dev)),
                                                  _p(linspace_table(w, dev)), _p(out), n, c, d, h, w, _FLAVOR["mode"], _stream()))
    return out


def warp3d_gather(tenInput: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    """warp3d on the global-gather kernel only (ofsv_warp3d_gather_f32); `warp3d` itself picks the TMA slab kernel on cubic
    volumes.  Bit-identical results; for tests and benchmarks."""
    x, f = _cuda_f32(tenInput, "tenInput"), _cuda_f32(tenFlow, "tenFlow")
    _check_warp_shapes(x, f, 3)
    n, c, d, h, w = x.shape
    dev = x.device
    out = torch.empty_like(x)
    with _on(dev):
        _C.check(_C.lib().ofsv_warp3d_gather_f32(_p(x), _p(f), _p(linspace_table(h, dev)), _p(linspace_table(d, dev)),
                                                 _p(linspace_table(w, dev)), _p(out), n, c, d, h, w, _FLAVOR["mode"], _stream()))
    return out


def warp_bwd(tenInput, tenFlow, grad_out, need_input_grad=True, need_flow_grad=True):
    """Backward of `warp` (ofsv_warp{2,3}d_bwd_f32): (grad_input, grad_flow), either None when not asked for.
    = ATen grid_sampler backward (bilinear / border / align_corners, zero gradient where the coordinate was clipped)
    chained with the backward of flow / ((S-1)/2) — what autograd runs under Flow-*/model/warplayer.py:26 / :37."""
    x, f, go = _cuda_f32(tenInput, "tenInput"), _cuda_f32(tenFlow, "tenFlow"), _cuda_f32(grad_out, "grad_out")
    nd = x.dim() - 2
    _check_warp_shapes(x, f, nd)
    if go.shape != x.shape:
        raise ValueError(f"warp_bwd: grad_out {tuple(go.shape)} does not match the output shape {tuple(x.shape)}")
    dev = x.device
    gx = torch.empty_like(x) if need_input_grad else None       # zero-filled by the library
    gf = torch.empty_like(f) if need_flow_grad else None
    with _on(dev), _span("warp_bwd"):
        if nd == 2:
            n, c, h, w = x.shape
            _C.check(_C.lib().ofsv_warp2d_bwd_f32(_p(x), _p(f), _p(go), _p(linspace_table(w, dev)), _p(linspace_table(h, dev)),
                                                  _p(gx), _p(gf), n, c, h, w, _FLAVOR["mode"], _stream()))
        else:
            n, c, d, h, w = x.shape
            _C.check(_C.lib().ofsv_warp3d_bwd_f32(_p(x), _p(f), _p(go), _p(linspace_table(h, dev)), _p(linspace_table(d, dev)),
                                                  _p(linspace_table(w, dev)), _p(gx), _p(gf), n, c, d, h, w,
                                                  _FLAVOR["mode"], _stream()))
    return gx, gf


class _WarpFn(torch.autograd.Function):
    """autograd node of warp(): the reference's warp is differentiable in both arguments (it is a grid_sample)."""

    @staticmethod
    def forward(ctx, x, f):
        ctx.save_for_backward(x, f)
        return _warp_fwd(x, f)

    @staticmethod
    def backward(ctx, grad_out):
        x, f = ctx.saved_tensors
        gx, gf = warp_bwd(x, f, grad_out, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gx, gf


def _warp(tenInput, tenFlow, nd):
    x, f = _cuda_f32(tenInput, "tenInput"), _cuda_f32(tenFlow, "tenFlow")
    _check_warp_shapes(x, f, nd)
    if torch.is_grad_enabled() and (x.requires_grad or f.requires_grad):
        return _WarpFn.apply(x, f)
    return _warp_fwd(x, f)


def warp2d(tenInput: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    return _warp(tenInput, tenFlow, 2)


def warp3d(tenInput: torch.Tensor, tenFlow: torch.Tensor) -> torch.Tensor:
    return _warp(tenInput, tenFlow, 3)


def warp_blend(img0, img1, flow, mask_logit, want_warped=True, want_merged=True, want_mask=True):
    """sigmoid(mask); warp(img0, flow[:, :nd]); warp(img1, flow[:, nd:]); merged = w0*m + w1*(1-m) in one pass.
    Returns (warped0, warped1, merged, mask_sigmoid); entries not asked for are None."""
    img0, img1, flow = _cuda_f32(img0, "img0"), _cuda_f32(img1, "img1"), _cuda_f32(flow, "flow")
    nd = img0.dim() - 2
    if nd not in (2, 3) or img0.shape[1] != 1 or img1.shape != img0.shape or flow.shape[1] != 2 * nd \
            or flow.shape[2:] != img0.shape[2:] or flow.shape[0] != img0.shape[0]:
        raise ValueError("warp_blend: bad shapes")
    need_m = want_merged or want_mask
    if need_m:
        mask_logit = _cuda_f32(mask_logit, "mask_logit")
        if mask_logit.shape != img0.shape:
            raise ValueError("warp_blend: mask shape")
    # on cubic 3-D volumes the library runs two TMA slab warps + a blend pass when it is given both warped buffers (35 % faster
    # than its single gather kernel): hand it scratch buffers even when the caller does not want the warped volumes
    slab = nd == 3 and img0.shape[2] == img0.shape[3] == img0.shape[4] and img0.shape[2] % 32 == 0
    w0 = torch.empty_like(img0) if (want_warped or slab) else None
    w1 = torch.empty_like(img0) if (want_warped or slab) else None
    mg = torch.empty_like(img0) if want_merged else None
    ms = torch.empty_like(img0) if want_mask else None
    dev = img0.device
    L = _C.lib()
    with _on(dev), _span("warp_blend"):
        if nd == 2:
            n, _, h, w = img0.shape
            _C.check(L.ofsv_warp_blend_2d_f32(_p(img0), _p(img1), _p(flow), _p(mask_logit if need_m else None),
                                              _p(linspace_table(w, dev)), _p(linspace_table(h, dev)), _p(w0), _p(w1), _p(mg),
                                              _p(ms), n, h, w, _FLAVOR["mode"], _stream()))
        else:
            n, _, d, h, w = img0.shape
            _C.check(L.ofsv_warp_blend_3d_f32(_p(img0), _p(img1), _p(flow), _p(mask_logit if need_m else None),
                                              _p(linspace_table(h, dev)), _p(linspace_table(d, dev)), _p(linspace_table(w, dev)),
                                              _p(w0), _p(w1), _p(mg), _p(ms), n, d, h, w, _FLAVOR["mode"], _stream()))
    if not want_warped:
        w0 = w1 = None
    return w0, w1, mg, ms


def blend(w0, w1, mask_logit):
    w0, w1, m = _cuda_f32(w0, "w0"), _cuda_f32(w1, "w1"), _cuda_f32(mask_logit, "mask_logit")
    if not (w0.shape == w1.shape == m.shape):
        raise ValueError("blend: shapes differ")
    out = torch.empty_like(w0)
    with _on(w0.device):
        _C.check(_C.lib().ofsv_blend_f32(_p(w0), _p(w1), _p(m), _p(out), w0.numel(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ UPFlow ops
def corr81_fwd(f1, f2, leaky_slope=None, out=None):
    """(B,C,H,W) x2 -> (B,81,H,W).  `out` may be a (B,>=81,H,W) channel-slice view of a larger contiguous buffer
    (the estimator's concat input): only out[:, :81] is written."""
    f1, f2 = _cuda_f32(f1, "input1"), _cuda_f32(f2, "input2")
    if f1.dim() != 4 or f1.shape != f2.shape:
        raise ValueError(f"corr81: bad shapes {tuple(f1.shape)} / {tuple(f2.shape)}")
    b, c, h, w = f1.shape
    if out is None:
        out = torch.empty(b, 81, h, w, device=f1.device, dtype=torch.float32)
    else:
        if (not out.is_cuda or out.dtype != torch.float32 or out.dim() != 4 or out.shape[0] != b or out.shape[1] < 81
                or out.shape[2:] != f1.shape[2:] or out.stride()[1:] != (h * w, w, 1)):
            raise ValueError("corr81: `out` must be a float32 CUDA (B,>=81,H,W) tensor with dense (C,H,W) strides")
    bstride = out.stride(0) if b > 1 else 81 * h * w
    with _on(f1.device):
        L = _C.lib()
        ns = L.ofsv_corr81_fwd_splits(b, c, h, w)          # coarse pyramid levels: channels split over CTAs through a scratch buffer
        work = torch.empty((ns, b, 81, h, w), device=f1.device, dtype=torch.float32) if ns > 1 else None
        _C.check(L.ofsv_corr81_fwd_f32(_p(f1), _p(f2), _p(out), b, c, h, w,
                                       float(leaky_slope or 0.0), int(leaky_slope is not None),
                                       max(bstride, 81 * h * w), _p(work), _stream()))
    return out


def corr81_bwd(f1, f2, gout):
    f1, f2, gout = _cuda_f32(f1, "input1"), _cuda_f32(f2, "input2"), _cuda_f32(gout, "grad_output")
    b, c, h, w = f1.shape
    if gout.shape != (b, 81, h, w):
        raise ValueError("corr81_bwd: grad_output shape")
    g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
    with _on(f1.device):
        _C.check(_C.lib().ofsv_corr81_bwd_f32(_p(f1), _p(f2), _p(gout), _p(g1), _p(g2), b, c, h, w, _stream()))
    return g1, g2


def _upsample_flow_ac_fwd(flow, h, w, if_rate):
    b, _, h_, w_ = flow.shape
    out = torch.empty(b, 2, h, w, device=flow.device, dtype=torch.float32)
    with _on(flow.device):
        _C.check(_C.lib().ofsv_upsample_flow_ac_f32(_p(flow), _p(out), b, h_, w_, h, w, int(if_rate), _stream()))
    return out


def upsample_flow_ac_bwd(grad_out, h_in, w_in, if_rate=True):
    """Backward of upsample2d_flow_as (ofsv_upsample_flow_ac_bwd_f32): grad wrt the (B,2,h_in,w_in) input."""
    go = _cuda_f32(grad_out, "grad_out")
    if go.dim() != 4 or go.shape[1] != 2:
        raise ValueError("upsample_flow_ac_bwd: expected (B,2,h,w)")
    b, _, h, w = go.shape
    gin = torch.empty(b, 2, h_in, w_in, device=go.device, dtype=torch.float32)     # zero-filled by the library
    with _on(go.device):
        _C.check(_C.lib().ofsv_upsample_flow_ac_bwd_f32(_p(go), _p(gin), b, h_in, w_in, h, w, int(if_rate), _stream()))
    return gin


class _UpsampleFlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, h, w, if_rate):
        ctx.meta = (flow.shape[2], flow.shape[3], if_rate)
        return _upsample_flow_ac_fwd(flow, h, w, if_rate)

    @staticmethod
    def backward(ctx, grad_out):
        h_, w_, if_rate = ctx.meta
        return upsample_flow_ac_bwd(grad_out, h_, w_, if_rate), None, None, None


def upsample_flow_ac(flow, h, w, if_rate=True):
    flow = _cuda_f32(flow, "inputs")
    if flow.dim() != 4 or flow.shape[1] != 2:
        raise ValueError("upsample_flow_ac: expected (B,2,h,w)")
    if torch.is_grad_enabled() and flow.requires_grad:
        return _UpsampleFlowFn.apply(flow, h, w, if_rate)
    return _upsample_flow_ac_fwd(flow, h, w, if_rate)


def _warping_no_div_fwd(x, flow):
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    with _on(x.device):
        _C.check(_C.lib().ofsv_warping_no_div_f32(_p(x), _p(flow), _p(out), b, c, h, w, _FLAVOR["mode"], _stream()))
    return out


def warping_no_div_bwd(x, flow, grad_out, need_input_grad=True, need_flow_grad=True):
    """Backward of WarpingLayer_no_div (ofsv_warping_no_div_bwd_f32): (grad_x, grad_flow), either None when not asked for."""
    x, flow, go = _cuda_f32(x, "x"), _cuda_f32(flow, "flow"), _cuda_f32(grad_out, "grad_out")
    if x.dim() != 4 or flow.shape != (x.shape[0], 2, x.shape[2], x.shape[3]) or go.shape != x.shape:
        raise ValueError("warping_no_div_bwd: bad shapes")
    b, c, h, w = x.shape
    gx = torch.empty_like(x) if need_input_grad else None          # zero-filled by the library
    gf = torch.empty_like(flow) if need_flow_grad else None
    with _on(x.device):
        _C.check(_C.lib().ofsv_warping_no_div_bwd_f32(_p(x), _p(flow), _p(go), _p(gx), _p(gf), b, c, h, w, _FLAVOR["mode"], _stream()))
    return gx, gf


class _WarpingNoDivFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow):
        ctx.save_for_backward(x, flow)
        return _warping_no_div_fwd(x, flow)

    @staticmethod
    def backward(ctx, grad_out):
        x, flow = ctx.saved_tensors
        return warping_no_div_bwd(x, flow, grad_out, ctx.needs_input_grad[0], ctx.needs_input_grad[1])


def warping_no_div(x, flow):
    x, flow = _cuda_f32(x, "x"), _cuda_f32(flow, "flow")
    if x.dim() != 4 or flow.shape != (x.shape[0], 2, x.shape[2], x.shape[3]):
        raise ValueError("warping_no_div: bad shapes")
    if torch.is_grad_enabled() and (x.requires_grad or flow.requires_grad):
        return _WarpingNoDivFn.apply(x, flow)
    return _warping_no_div_fwd(x, flow)


def torch_warp(x, flow):
    """tools.torch_warp (UPFlow/utils/tools.py:1317-1361): WarpingLayer_no_div's sampling without the validity mask (forward only)."""
    x, flow = _cuda_f32(x, "x"), _cuda_f32(flow, "flow")
    if x.dim() != 4 or flow.shape != (x.shape[0], 2, x.shape[2], x.shape[3]):
        raise ValueError("torch_warp: bad shapes")
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    with _on(x.device), _span("torch_warp"):
        _C.check(_C.lib().ofsv_torch_warp_f32(_p(x), _p(flow), _p(out), b, c, h, w, _FLAVOR["mode"], _stream()))
    return out


def feature_norm_pair(f_plain, f_src, flow=None):
    """(normalize(f_plain), normalize(WarpingLayer_no_div(f_src, flow))) in one launch (ofsv_feature_norm_pair_f32) — the producer of
    the cost-volume inputs, UPFlow/model/upflow.py:621-640 with the per-(sample, channel) moments of simple_train.py:321-329.
    flow None = pyramid level 0 (no warp)."""
    f_plain, f_src = _cuda_f32(f_plain, "f_plain"), _cuda_f32(f_src, "f_src")
    if f_plain.dim() != 4 or f_src.shape != f_plain.shape:
        raise ValueError("feature_norm_pair: bad shapes")
    b, c, h, w = f_plain.shape
    if flow is not None:
        flow = _cuda_f32(flow, "flow")
        if flow.shape != (b, 2, h, w):
            raise ValueError("feature_norm_pair: bad flow shape")
    o0, o1 = torch.empty_like(f_plain), torch.empty_like(f_src)
    with _on(f_plain.device), _span("feature_norm_pair"):
        _C.check(_C.lib().ofsv_feature_norm_pair_f32(_p(f_plain), _p(f_src), _p(flow), _p(o0), _p(o1), b, c, h, w, _FLAVOR["mode"], _stream()))
    return o0, o1


# ------------------------------------------------------------------------------------------------ IFNet engine pieces
_TDT = {_C.F32: torch.float32, _C.BF16: torch.bfloat16}


_WS = {}


def workspace(key, shape, dtype, device):
    """Persistent ZERO-initialised buffer per (key, shape, dtype, device).  Used for the shifted space-to-depth tensors whose
    border sub-cells (the conv padding) must stay zero: producers only ever write the interior, so the buffer is zeroed
    once and reused by every later call on the same stream."""
    # one buffer per host thread: two threads driving models on their own streams must not share a workspace (within a thread the
    # calls are stream-ordered; running ONE thread's calls concurrently on several streams is not supported)
    k = (key, tuple(shape), dtype, str(device), threading.get_ident())
    t = _WS.get(k)
    if t is None:
        t = torch.zeros(shape, dtype=dtype, device=device)
        _WS[k] = t
    return t


def clear_workspaces():
    _WS.clear()


def s2d_shape(n, sp, c):
    """Shape of the shifted space-to-depth tensor of a logical [n][*sp][c] channels-last tensor (sp = (D,H,W) or (H,W))."""
    return [n] + [v // 2 + 1 for v in sp] + [(2 ** len(sp)) * c]


def pack_nhwc(srcs, cs):
    """Channel concatenation of 1..8 fp32 (N,C_i,*sp) CUDA tensors -> channels-last bf16 [N][D][H][W][cs] (D = 1 for 2-D inputs),
    zero-padded channels, in one launch (ofsv_pack_nhwc_bf16)."""
    srcs = [_cuda_f32(t, "pack_nhwc source") for t in srcs]
    n, sp = srcs[0].shape[0], tuple(srcs[0].shape[2:])
    if not 1 <= len(srcs) <= 8 or any(t.shape[0] != n or tuple(t.shape[2:]) != sp for t in srcs):
        raise ValueError("pack_nhwc: 1..8 tensors with equal batch and spatial dims")
    pix = 1
    for v in sp:
        pix *= v
    out = torch.empty((n,) + ((1,) + sp if len(sp) == 2 else sp) + (cs,), device=srcs[0].device, dtype=torch.bfloat16)
    ptrs = (ctypes.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
    chans = (ctypes.c_int * len(srcs))(*[t.shape[1] for t in srcs])
    with _on(out.device), _span("pack_nhwc"):
        _C.check(_C.lib().ofsv_pack_nhwc_bf16(ptrs, chans, len(srcs), _p(out), n, pix, cs, _stream()))
    return out


def unpack_nhwc(y, c, nd):
    """Channels-last bf16 [N][D][H][W][Cs] (D = 1 for nd = 2) -> fp32 (N,c,*sp) with the first c channels (ofsv_unpack_nhwc_f32)."""
    if not (y.is_cuda and y.dtype == torch.bfloat16 and y.is_contiguous() and y.dim() == 5):
        raise TypeError("unpack_nhwc: expected a contiguous CUDA bf16 [N][D][H][W][Cs] tensor (no CPU path)")
    n, d, h, w, cs = y.shape
    out = torch.empty((n, c) + ((h, w) if nd == 2 else (d, h, w)), device=y.device, dtype=torch.float32)
    with _on(y.device), _span("unpack_nhwc"):
        _C.check(_C.lib().ofsv_unpack_nhwc_f32(_p(y), _p(out), n, d * h * w, cs, c, _stream()))
    return out


def pack_block_input(img0, img1, warped0, warped1, mask, flow, scale, act_dtype, cs=16, s2d=False, key="xin"):
    nd = img0.dim() - 2
    n = img0.shape[0]
    sp = list(img0.shape[2:])
    d, h, w = ([1] + sp) if nd == 2 else sp
    osp = [s // scale for s in sp]
    if s2d:
        dst = workspace((key, scale), s2d_shape(n, osp, cs), _TDT[act_dtype], img0.device)
    else:
        dst = torch.empty([n] + osp + [cs], device=img0.device, dtype=_TDT[act_dtype])
    with _on(img0.device), _span("pack_block_input"):
        _C.check(_C.lib().ofsv_pack_block_input(_p(img0), _p(img1), _p(warped0), _p(warped1), _p(mask), _p(flow), _p(dst),
                                                act_dtype, nd, n, d, h, w, scale, cs, int(s2d), _stream()))
    return dst


def head_upsample_add(head, flow_prev, mask_prev, nd, n, sp, scale):
    """head [N][sp/scale...][Cs] fp32 -> (flow (N,2nd,*sp), mask (N,1,*sp))."""
    d, h, w = ([1] + list(sp)) if nd == 2 else list(sp)
    flow = torch.empty([n, 2 * nd] + list(sp), device=head.device, dtype=torch.float32)
    mask = torch.empty([n, 1] + list(sp), device=head.device, dtype=torch.float32)
    with _on(head.device), _span("head_upsample_add"):
        _C.check(_C.lib().ofsv_head_upsample_add(_p(head), head.shape[-1], _p(flow_prev), _p(mask_prev), _p(flow), _p(mask),
                                                 nd, n, d, h, w, scale, _stream()))
    return flow, mask


def head_upsample_add_bwd(gflow, gmask, nd, scale):
    """Gradient of head_upsample_add w.r.t. the head: channels-last fp32 [N][D/s][H/s][W/s][8] (ofsv_head_upsample_add_bwd)."""
    gflow, gmask = _cuda_f32(gflow, "gflow"), _cuda_f32(gmask, "gmask")
    n = gflow.shape[0]
    sp = tuple(gflow.shape[2:])
    d, h, w = ((1,) + sp) if nd == 2 else sp
    out = torch.empty((n, (d // scale) if nd == 3 else 1, h // scale, w // scale, 8), device=gflow.device, dtype=torch.float32)
    with _on(gflow.device), _span("head_upsample_add_bwd"):
        _C.check(_C.lib().ofsv_head_upsample_add_bwd(_p(gflow), _p(gmask), _p(out), nd, n, d, h, w, scale, _stream()))
    return out


def pack_block_input_bwd(gx, nd, sp, scale):
    """Gradients of pack_block_input (plain layout, Cs = 16) w.r.t. (warped0, warped1, mask, flow) (ofsv_pack_block_input_bwd)."""
    if not (gx.is_cuda and gx.dtype == torch.bfloat16 and gx.is_contiguous() and gx.shape[-1] == 16):
        raise TypeError("pack_block_input_bwd: expected a contiguous CUDA bf16 [...][16] tensor (no CPU path)")
    n = gx.shape[0]
    d, h, w = ((1,) + tuple(sp)) if nd == 2 else tuple(sp)
    g0 = torch.empty((n, 1) + tuple(sp), device=gx.device, dtype=torch.float32)
    g1, gm = torch.empty_like(g0), torch.empty_like(g0)
    gf = torch.empty((n, 2 * nd) + tuple(sp), device=gx.device, dtype=torch.float32)
    with _on(gx.device), _span("pack_block_input_bwd"):
        _C.check(_C.lib().ofsv_pack_block_input_bwd(_p(gx), _p(g0), _p(g1), _p(gm), _p(gf), nd, n, d, h, w, scale, _stream()))
    return g0, g1, gm, gf


def block_stage_3d(head, fm_prev, img0, img1, scale_head, scale_next, want_merged, want_mask, pack_s2d=False, key="xin", hfast=False):
    """Fused 3-D block output stage on the channels-last state (ofsv_block_stage_3d).  State layout: [N,D,H,W,8] fp32, or with
    `hfast` the H-fastest [N,D,W,H,8] (_C.STATE_DWH8) that csrc/block_stage_hfast.cu works on; `head` is the block's head at
    1/scale_head resolution in the same layout, fm_prev the previous state or None.  scale_head = 0: fm_prev already holds the
    accumulated state (the head conv's epilogue added the head), only warp / blend / pack run.
    Returns (fm, merged|None, mask_sig|None, next_block_input|None)."""
    n, _, d, h, w = img0.shape
    dev = img0.device
    st_shape = (n, d, w, h, 8) if hfast else (n, d, h, w, 8)
    if scale_head == 0:
        if head is not None or fm_prev is None:
            raise ValueError("block_stage_3d: scale_head = 0 takes the already accumulated state in fm_prev and no head")
    else:
        s = scale_head
        hd_shape = (n, d // s, w // s, h // s, 8) if hfast else (n, d // s, h // s, w // s, 8)
        if head.dtype != torch.float32 or tuple(head.shape) != hd_shape or not head.is_contiguous():
            raise ValueError(f"block_stage_3d: head must be a contiguous fp32 {hd_shape} tensor, got {tuple(head.shape)}")
    if fm_prev is not None and (tuple(fm_prev.shape) != st_shape or fm_prev.dtype != torch.float32 or not fm_prev.is_contiguous()):
        raise ValueError(f"block_stage_3d: fm_prev must be a contiguous fp32 {st_shape} tensor")
    fm = torch.empty(st_shape, device=dev, dtype=torch.float32) if scale_head else None
    mg = torch.empty((n, 1, d, h, w), device=dev, dtype=torch.float32) if want_merged else None
    ms = torch.empty((n, 1, d, h, w), device=dev, dtype=torch.float32) if want_mask else None
    pk = None
    if scale_next:
        osp = [d // scale_next, h // scale_next, w // scale_next]
        if pack_s2d:
            pk = workspace((key, scale_next), s2d_shape(n, osp, 16), torch.bfloat16, dev)
        else:
            pk = torch.empty([n] + osp + [16], device=dev, dtype=torch.bfloat16)
    with _on(dev), _span("block_stage"):
        _C.check(_C.lib().ofsv_block_stage_3d(_p(head), _p(fm_prev), _p(img0), _p(img1), _p(linspace_table(h, dev)),
                                              _p(linspace_table(d, dev)), _p(linspace_table(w, dev)), _p(fm), _p(mg), _p(ms),
                                              _p(pk), n, d, h, w, scale_head, scale_next, int(bool(pack_s2d) and scale_next != 0), _FLAVOR["mode"],
                                              _C.STATE_DWH8 if hfast else _C.STATE_DHW8, _stream()))
    return (fm if scale_head else fm_prev), mg, ms, pk


def state_views(fm, hfast=False):
    """(flow (N,6,D,H,W), mask_logit (N,1,D,H,W)) as permuted VIEWS of the channels-last state ([N,D,H,W,8], or the H-fastest
    [N,D,W,H,8] with `hfast`)."""
    perm = (0, 4, 1, 3, 2) if hfast else (0, 4, 1, 2, 3)
    return fm[..., :6].permute(*perm), fm[..., 6:7].permute(*perm)


def conv_pack_weights(desc: "_C.ConvDesc", w_tap: torch.Tensor, layout: int) -> torch.Tensor:
    """fp32 tap-form weights [T][Cin_s][Cout_w] -> the bf16 K-major blocks of `layout` (_C.WL_TAP | _C.WL_STACK) in ONE launch
    (ofsv_conv_pack_weights).  Only the layer structure of `desc` is read (taps, Cin_s, Cout_w)."""
    w = _cuda_f32(w_tap, "w_tap")
    out = torch.empty(w.numel(), dtype=torch.bfloat16, device=w.device)
    with _on(w.device):
        _C.check(_C.lib().ofsv_conv_pack_weights(ctypes.byref(desc), _p(w), _p(out), int(layout), _stream()))
    return out


def conv_halo_weight_layout(desc: "_C.ConvDesc") -> int:
    rc = _C.lib().ofsv_conv_halo_weight_layout(ctypes.byref(desc))
    if rc < 0:
        _C.check(rc)
    return rc


def set_tuning(key: str, value: int) -> None:
    """Process-wide A/B switch between equivalent code paths of the library (ofsv_set_tuning)."""
    _C.check(_C.lib().ofsv_set_tuning(key.encode(), int(value)))


def conv(desc: "_C.ConvDesc", x, w, bias, prelu, residual, y, engine: str):
    L = _C.lib()
    fn = {"tc": L.ofsv_conv_tc, "halo": L.ofsv_conv_halo, "simt": L.ofsv_conv_simt}[engine]
    with _on(x.device), _span("conv_" + engine):
        _C.check(fn(ctypes.byref(desc), _p(x), _p(w), _p(bias), _p(prelu), _p(residual), _p(y), _stream()))
    return y


def u8_to_f32(src: torch.Tensor, div: float = 255.0) -> torch.Tensor:
    """uint8 CUDA tensor -> fp32 `src / div` (IEEE division, bit-identical to `src.float() / div`) in one pass."""
    if not isinstance(src, torch.Tensor) or not src.is_cuda or src.dtype != torch.uint8:
        raise TypeError("u8_to_f32: expected a uint8 CUDA tensor")
    src = src.contiguous()
    dst = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    with _on(src.device):
        _C.check(_C.lib().ofsv_u8_to_f32(_p(src), _p(dst), src.numel(), float(div), _stream()))
    return dst


METRIC_BLOCKS = 64


def f32_to_u8(src: torch.Tensor, mul: float = 255.0) -> torch.Tensor:
    """(src * mul).byte() with the product clamped to [0, 255] (ofsv_f32_to_u8): the reference's export
    `(img * 255).byte()` (Flow-3D/inference_img.py:105) on the device, so that byte volumes are downloaded as bytes."""
    x = _cuda_f32(src, "src")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with _on(x.device):
        _C.check(_C.lib().ofsv_f32_to_u8(_p(x), _p(out), x.numel(), float(mul), _stream()))
    return out


def _metric_args(a: torch.Tensor, b: torch.Tensor, what: str):
    if not (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.is_cuda and b.is_cuda):
        raise TypeError(f"{what}: expected CUDA tensors (there is no CPU path)")
    if a.shape != b.shape:
        raise ValueError(f"{what}: Input images must have the same dimensions.")
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32 tensors")
    return a.contiguous(), b.contiguous()


def sq_err_sums(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Per-sample sum of ((a - b) * scale)^2 over all but the first dimension -> float64 [N] (deterministic)."""
    a, b = _metric_args(a, b, "sq_err_sums")
    n = a.shape[0] if a.dim() > 0 else 1
    count = a.numel() // max(n, 1)
    out = torch.empty(n, dtype=torch.float64, device=a.device)
    part = torch.empty(max(n, 1) * METRIC_BLOCKS, dtype=torch.float64, device=a.device)
    with _on(a.device), _span("sq_err"):
        _C.check(_C.lib().ofsv_sq_err_f64(_p(a), _p(b), _p(part), _p(out), n, count, float(scale), _stream()))
    return out


def ssim2d_means(x: torch.Tensor, y: torch.Tensor, data_range: float = 255.0) -> torch.Tensor:
    """Mean SSIM (11x11 Gaussian window, sigma 1.5, valid region) of every [H][W] plane pair of x, y [..., H, W] -> float64."""
    x, y = _metric_args(x, y, "ssim2d_means")
    if x.dim() < 2:
        raise ValueError("ssim2d_means: Wrong input image dimensions.")
    H, W = int(x.shape[-2]), int(x.shape[-1])
    n = x.numel() // (H * W)
    out = torch.empty(n, dtype=torch.float64, device=x.device)
    part = torch.empty(max(n, 1) * METRIC_BLOCKS, dtype=torch.float64, device=x.device)
    with _on(x.device), _span("ssim2d"):
        _C.check(_C.lib().ofsv_ssim2d_f64(_p(x), _p(y), _p(part), _p(out), n, H, W, float(data_range), _stream()))
    return out.reshape(x.shape[:-2])


def launch_count() -> int:
    return int(_C.lib().ofsv_launch_count())
