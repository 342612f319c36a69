"""Micro-benchmark of the UPFlow correlation cost volume (a8) on the five pyramid shapes of a 256x832 pair,
B = 16 (8 pairs x 2 directions): GB/s on the algorithmic bytes (2C + 81) * H * W * 4 per sample."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops
B = 16
tot_b = tot_t = 0.0
for (c, h, w) in ((196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)):
    f1, f2 = torch.randn(B, c, h, w, device="cuda"), torch.randn(B, c, h, w, device="cuda")
    g = torch.randn(B, 81, h, w, device="cuda")
    for name, fn, nbytes in (("fwd", lambda: ops.corr81_fwd(f1, f2, leaky_slope=0.1), (2 * c + 81) * h * w * 4 * B),
                             ("bwd", lambda: ops.corr81_bwd(f1, f2, g), (4 * c + 81) * h * w * 4 * B)):
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print(f"corr81 {name} B={B} C={c:3d} {h:2d}x{w:3d}: {us:7.1f} us  {nbytes / us / 1e3:7.1f} GB/s algorithmic")
        if name == "fwd": tot_b += nbytes; tot_t += us
print(f"all five levels fwd: {tot_t:.1f} us for {tot_b / 1e6:.2f} MB -> {tot_b / tot_t / 1e3:.1f} GB/s (launch-bound: 5 launches)")
