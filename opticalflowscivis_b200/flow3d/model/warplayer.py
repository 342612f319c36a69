"""Flow-3D/model/warplayer.py — `warp(tenInput, tenFlow)` on the sm_100a gather kernel (ofsv_warp3d_f32)."""
from ...ops import warp3d as _warp


def warp(tenInput, tenFlow):
    return _warp(tenInput, tenFlow)
