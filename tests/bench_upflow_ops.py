"""Micro-benchmark of the UPFlow flow-path operators, forward AND backward (a8, a10, a11 + their autograd), on the five pyramid
shapes of a 256x832 pair, B = 16 (8 pairs x 2 directions), plus the 3-D / 2-D warp backward at the BASELINE
volume / frame sizes.  Prints microseconds per call and GB/s on the algorithmic bytes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops


def timeit(fn, reps=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B = 16
tot = {}
for (c, h, w) in ((196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)):
    x, fl = torch.randn(B, c, h, w, device="cuda"), torch.randn(B, 2, h, w, device="cuda") * 2
    f2 = torch.randn(B, c, h, w, device="cuda")
    go, g81 = torch.randn(B, c, h, w, device="cuda"), torch.randn(B, 81, h, w, device="cuda")
    gup = torch.randn(B, 2, 2 * h, 2 * w, device="cuda")
    cases = (("warping_no_div fwd", lambda: ops.warping_no_div(x, fl), (2 * c + 2) * h * w * 4 * B),
             ("warping_no_div bwd", lambda: ops.warping_no_div_bwd(x, fl, go), (3 * c + 4) * h * w * 4 * B),
             ("corr81 fwd", lambda: ops.corr81_fwd(x, f2, leaky_slope=0.1), (2 * c + 81) * h * w * 4 * B),
             ("corr81 bwd", lambda: ops.corr81_bwd(x, f2, g81), (4 * c + 81) * h * w * 4 * B),
             ("upsample_flow x2 fwd", lambda: ops.upsample_flow_ac(fl, 2 * h, 2 * w), (2 + 8) * h * w * 4 * B),
             ("upsample_flow x2 bwd", lambda: ops.upsample_flow_ac_bwd(gup, h, w), (2 + 8) * h * w * 4 * B))
    for name, fn, nbytes in cases:
        us = timeit(fn)
        tot[name] = tot.get(name, 0.0) + us
        print(f"{name:22s} B={B} C={c:3d} {h:2d}x{w:3d}: {us:7.1f} us  {nbytes / us / 1e3:7.1f} GB/s algorithmic")
print("sum over the five levels (us): " + ", ".join(f"{k} {v:.0f}" for k, v in tot.items()))

# warp backward at the BASELINE sizes: bytes = flow + gout + src read, gsrc (memset + red) + gflow written
for name, shape in (("warp3d bwd 256^3", (1, 1, 256, 256, 256)), ("warp3d bwd 4x128^3", (4, 1, 128, 128, 128)), ("warp2d bwd 64x160x224", (64, 1, 160, 224))):
    nd = len(shape) - 2
    src = torch.rand(shape, device="cuda")
    small = tuple(max(1, s // 8) for s in shape[2:])
    fl = torch.nn.functional.interpolate(torch.randn((shape[0], nd) + small, device="cuda") * 2, size=shape[2:],
                                         mode="trilinear" if nd == 3 else "bilinear").contiguous()
    go = torch.randn(shape, device="cuda")
    vox = src.numel()
    us = timeit(lambda: ops.warp_bwd(src, fl, go), 20)
    nbytes = (nd + 1 + 1 + 1 + nd) * 4 * vox
    print(f"{name:22s}: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s algorithmic ({(2 * nd + 3) * 4} B/voxel); forward: {timeit(lambda: (ops.warp3d if nd == 3 else ops.warp2d)(src, fl), 20):.1f} us")
