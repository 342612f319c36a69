"""ctypes/numpy binding of oracle/libofsv_oracle.so.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DIV_TRUE, DIV_RCP = 0, 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libofsv_oracle.so")
    src = os.path.join(_HERE, "ofsv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libofsv_oracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _lin(n):
    import torch
    return torch.linspace(-1.0, 1.0, n).numpy()


def warp2d(src, flow, div_mode=DIV_TRUE, lin_x=None, lin_y=None):
    src, ps = _f(src); flow, pf = _f(flow)
    n, c, h, w = src.shape
    lx, plx = _f(_lin(w) if lin_x is None else lin_x); ly, ply = _f(_lin(h) if lin_y is None else lin_y)
    out = np.empty_like(src)
    assert lib().ofsv_oracle_warp2d(ps, pf, plx, ply, out.ctypes.data_as(ctypes.c_void_p), n, c, h, w, div_mode) == 0
    return out


def warp3d(src, flow, div_mode=DIV_TRUE, lins=None):
    src, ps = _f(src); flow, pf = _f(flow)
    n, c, d, h, w = src.shape
    lh, ld, lw = lins if lins is not None else (_lin(h), _lin(d), _lin(w))
    lh, plh = _f(lh); ld, pld = _f(ld); lw, plw = _f(lw)
    out = np.empty_like(src)
    assert lib().ofsv_oracle_warp3d(ps, pf, plh, pld, plw, out.ctypes.data_as(ctypes.c_void_p), n, c, d, h, w, div_mode) == 0
    return out


def blend(w0, w1, mask_logit):
    w0, p0 = _f(w0); w1, p1 = _f(w1); m, pm = _f(mask_logit)
    out = np.empty_like(w0)
    assert lib().ofsv_oracle_blend(p0, p1, pm, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(w0.size)) == 0
    return out


def corr81(f1, f2, leaky_slope=None):
    f1, p1 = _f(f1); f2, p2 = _f(f2)
    b, c, h, w = f1.shape
    out = np.empty((b, 81, h, w), np.float32)
    assert lib().ofsv_oracle_corr81(p1, p2, out.ctypes.data_as(ctypes.c_void_p), b, c, h, w,
                                    ctypes.c_float(leaky_slope or 0.0), int(leaky_slope is not None)) == 0
    return out


def corr81_bwd(f1, f2, gout):
    f1, p1 = _f(f1); f2, p2 = _f(f2); gout, pg = _f(gout)
    b, c, h, w = f1.shape
    g1 = np.empty_like(f1); g2 = np.empty_like(f2)
    assert lib().ofsv_oracle_corr81_bwd(p1, p2, pg, g1.ctypes.data_as(ctypes.c_void_p),
                                        g2.ctypes.data_as(ctypes.c_void_p), b, c, h, w) == 0
    return g1, g2


def upsample_flow_ac(flow, h, w, if_rate=True):
    flow, pf = _f(flow)
    b, two, h_, w_ = flow.shape
    assert two == 2
    out = np.empty((b, 2, h, w), np.float32)
    assert lib().ofsv_oracle_upsample_flow_ac(pf, out.ctypes.data_as(ctypes.c_void_p), b, h_, w_, h, w, int(if_rate)) == 0
    return out


def warping_no_div(x, flow, div_mode=DIV_TRUE):
    x, px = _f(x); flow, pf = _f(flow)
    b, c, h, w = x.shape
    out = np.empty_like(x)
    assert lib().ofsv_oracle_warping_no_div(px, pf, out.ctypes.data_as(ctypes.c_void_p), b, c, h, w, div_mode) == 0
    return out
