"""Fixed vs per-tile cost of one conv launch on the stacked tcgen05 engine: a 3^3 convblock layer (C -> C, residual) at growing
spatial sizes.  Prints microseconds per launch (CUDA events, 20 launches) and the engine's configuration."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import _C, ifnet, ops

dev = torch.device("cuda")
L = _C.lib()
buf = ctypes.create_string_buffer(512)
for c in (128, 64):
    blk = ifnet.IFBlock(3, 11, c).to(dev)
    lay = blk.layers()[3]                       # convblock0.1: C -> C, 27 taps, residual
    for n, s in ((1, 4), (1, 8), (1, 16), (4, 16), (1, 32), (4, 32), (1, 64)):
        sp = (s, s, s)
        d, osp = lay.desc(n, sp, _C.BF16)
        x = torch.randn((n,) + sp + (lay.cin_s,), device=dev).bfloat16()
        res = torch.randn((n,) + sp + (lay.cout_s,), device=dev).bfloat16()
        y = torch.empty((n,) + sp + (lay.cout_s,), device=dev, dtype=torch.bfloat16)
        w = lay.w_halo
        for _ in range(3):
            ops.conv(d, x, w, lay.bias, lay.prelu, res, y, "halo")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv(d, x, w, lay.bias, lay.prelu, res, y, "halo")
        e1.record(); torch.cuda.synchronize()
        L.ofsv_conv_halo_describe(ctypes.byref(d), buf, 512)
        tiles = n * (s // 16 if s >= 16 else 1) * (s // 8 if s >= 8 else 1) * s
        print(f"C={c:3d} N={n} {s:2d}^3: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us  (16x8 tiles x planes: {tiles:5d})  {buf.value.decode()}")
