"""ncu driver for the UPFlow correlation cost volume (a8): one forward (+LeakyReLU) and one backward launch pair per pyramid level of a
256x832 pair at B = 16, inside a cudaProfilerStart/Stop range (run ncu with --profile-from-start off)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops
B = 16
lv = []
for (c, h, w) in ((196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)):
    lv.append((torch.randn(B, c, h, w, device="cuda"), torch.randn(B, c, h, w, device="cuda"), torch.randn(B, 81, h, w, device="cuda")))
for _ in range(3):
    for f1, f2, g in lv:
        ops.corr81_fwd(f1, f2, leaky_slope=0.1); ops.corr81_bwd(f1, f2, g)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f1, f2, g in lv:
    ops.corr81_fwd(f1, f2, leaky_slope=0.1); ops.corr81_bwd(f1, f2, g)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
