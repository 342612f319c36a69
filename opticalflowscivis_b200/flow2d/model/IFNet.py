"""Flow-2D/model/IFNet.py — IFBlock / IFNet with the reference's constructor signatures."""
from ... import ifnet as _g
from .warplayer import warp            # noqa: F401  (the reference module re-exports it)


class IFBlock(_g.IFBlock):
    def __init__(self, in_planes, c=64):
        super().__init__(2, in_planes, c)


refine = False      # Flow-2D/model/IFNet.py:32 — module-level switch of the reference (read at construction)


class IFNet(_g.IFNet):
    def __init__(self, precision="bf16", engine="auto"):
        super().__init__(2, precision=precision, engine=engine, refine=refine)
