"""Flow-3D/model/RIFE.py — `Model` (inference surface; `update` is the next tier)."""
from ...rife import Model3D as Model   # noqa: F401
