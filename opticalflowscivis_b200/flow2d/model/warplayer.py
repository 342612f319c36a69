"""Flow-2D/model/warplayer.py — `warp(tenInput, tenFlow)` on the sm_100a gather kernel (ofsv_warp2d_f32)."""
from ...ops import warp2d as _warp


def warp(tenInput, tenFlow):
    return _warp(tenInput, tenFlow)
