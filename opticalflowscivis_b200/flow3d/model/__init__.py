from .warplayer import warp            # noqa: F401
from .IFNet import IFNet, IFBlock      # noqa: F401
from .RIFE import Model                # noqa: F401
