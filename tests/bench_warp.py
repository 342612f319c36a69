"""Micro-benchmark of the standalone warp kernels (a1/a2): GB/s on algorithmic bytes.  usage: bench_warp.py [size=256] [n=1] [flavor=cpu|cuda]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops, synth
s = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if len(sys.argv) > 3:
    ops.set_reference_flavor(sys.argv[3])
print('reference flavour:', ops.reference_flavor())
a, _, b = synth.droplet3d_u8(n, s)
src = torch.from_numpy(a).cuda().float() / 255
g = torch.Generator().manual_seed(7)
for kind in ("smooth", "zero", "noise"):
    if kind == "smooth":
        f = torch.nn.functional.interpolate((torch.randn((n, 3, s // 8, s // 8, s // 8), generator=g) * 2).cuda(), size=(s, s, s), mode="trilinear").contiguous()
    elif kind == "zero":
        f = torch.zeros((n, 3, s, s, s), device="cuda")
    else:
        f = (torch.randn((n, 3, s, s, s), generator=g) * 3).cuda()
    for _ in range(3):
        ops.warp3d(src, f)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.warp3d(src, f)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    for _ in range(3):
        ops.warp3d_gather(src, f)
    e0.record()
    for _ in range(20):
        ops.warp3d_gather(src, f)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 20
    print(f"warp3d {n}x{s}^3 {kind:6s}: {ms*1e3:8.1f} us  {20.0 * n * s**3 / ms / 1e6:8.1f} GB/s (20 B/voxel)   [global-gather kernel: {ms2*1e3:8.1f} us]")
    del f
# reference point: plain copy moving the same number of bytes
x = torch.empty(n * 5 * s**3 // 2, device="cuda"); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"copy of the same 20 B/voxel: {ms*1e3:8.1f} us  {20.0 * n * s**3 / ms / 1e6:8.1f} GB/s")
