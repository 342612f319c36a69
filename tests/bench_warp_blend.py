import os, sys, torch
sys.path.insert(0, "/root/repo")
from opticalflowscivis_b200 import ops
s = 256
img0, img1 = torch.rand(1, 1, s, s, s, device="cuda"), torch.rand(1, 1, s, s, s, device="cuda")
fl = torch.nn.functional.interpolate(torch.randn(1, 6, 32, 32, 32, device="cuda") * 2, size=(s, s, s), mode="trilinear").contiguous()
m = torch.randn(1, 1, s, s, s, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
us = t(lambda: ops.warp_blend(img0, img1, fl, m))
print(f"warp_blend 3d fused (7 in + 4 out planes = 52 B/voxel incl. 2 src): {us:.1f} us -> {52 * s**3 / us / 1e3:.0f} GB/s")
us2 = t(lambda: ops.warp_blend(img0, img1, fl, m, want_warped=False, want_mask=False))
print(f"warp_blend 3d merged only: {us2:.1f} us")
f0, f1 = fl[:, :3].contiguous(), fl[:, 3:].contiguous()
us3 = t(lambda: (ops.warp3d(img0, f0), ops.warp3d(img1, f1)))
print(f"two slab warps: {us3:.1f} us")
w0, w1 = ops.warp3d(img0, f0), ops.warp3d(img1, f1)
us4 = t(lambda: ops.blend(w0, w1, m))
print(f"blend: {us4:.1f} us")
