"""Optimizer step of the training tier on the 3-D IFNet parameter set: ofsv fused multi-tensor AdamW (one launch) vs
torch.optim.AdamW (foreach) and its fused=True variant.  Algorithmic bytes: 28 B per parameter (p, g, m, v read; p, m, v written)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200.optim import FusedAdamW, GradientBucket
from oracle.ifnet_ref import IFNetRef          # parameter container only (same shapes / keys as the reference's IFNet)


def timeit(fn, reps=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for nd in (2, 3):
    nets = [IFNetRef(nd).cuda() for _ in range(3)]
    n = sum(p.numel() for p in nets[0].parameters())
    for net in nets:
        for p in net.parameters():
            p.grad = torch.randn_like(p)
    bucket = GradientBucket(nets[0].parameters())
    bucket.flat.normal_()
    ours = FusedAdamW(nets[0].parameters(), bucket=bucket)
    t_foreach = torch.optim.AdamW(nets[1].parameters(), lr=1e-6, weight_decay=1e-3)
    t_fused = torch.optim.AdamW(nets[2].parameters(), lr=1e-6, weight_decay=1e-3, fused=True)
    a, b, c = timeit(ours.step), timeit(t_foreach.step), timeit(t_fused.step)
    print(f"IFNet{nd}D {n / 1e6:.2f} M parameters ({len(ours.params)} tensors): ofsv fused AdamW {a:7.1f} us = {28 * n / a / 1e3:6.0f} GB/s algorithmic; "
          f"torch foreach {b:7.1f} us; torch fused=True {c:7.1f} us")
