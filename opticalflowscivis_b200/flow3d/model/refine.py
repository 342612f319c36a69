"""Flow-3D/model/refine.py — Conv2 / Contextnet / Unet with the reference's constructor signatures."""
from ... import refine as _g
from .warplayer import warp            # noqa: F401  (the reference module imports it)

c = _g.C


class Conv2(_g.Conv2):
    def __init__(self, in_planes, out_planes, stride=2):
        super().__init__(3, in_planes, out_planes, stride)


class Contextnet(_g.Contextnet):
    def __init__(self):
        super().__init__(3)


class Unet(_g.Unet):
    def __init__(self):
        super().__init__(3)
