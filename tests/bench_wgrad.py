"""Per-layer timing of ofsv_conv_wgrad_bf16 on the layer shapes of one 8 x 64^3 training step (block2 / block_tea of the 3-D IFNet,
c = 64), brick-window kernel against the per-tap / tap-group kernels.  CUDA events, 20 launches each.  usage: bench_wgrad.py [n=8] [size=64]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import _C, ifnet, ops, train  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
blk = ifnet.IFBlock(3, 11, c=64).to(dev)
L = blk.layers()
cases = [("conv0.0", L[0], size), ("conv0.1", L[1], size // 2), ("convblock", L[2], size // 4), ("convT", L[10], size // 4), ("heads", L[11], size // 2)]
for name, lay, s in cases:
    d, osp = lay.desc(n, (s, s, s), _C.BF16, has_residual=False)
    x = torch.randn((n, s, s, s, lay.cin_s), device=dev).bfloat16()
    gy = torch.randn((n,) + tuple(osp) + (16 if lay.out_f32 else lay.cout_s,), device=dev).bfloat16()
    line = f"{name:10s} in {s}^3 x {lay.cin_s} -> {lay.cout_w}, {lay.nphase * lay.ntaps} taps:"
    for brick in (1, 0):
        ops.set_tuning("wgrad_brick", brick)
        for _ in range(3):
            train.conv_wgrad(d, x, gy)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            train.conv_wgrad(d, x, gy)
        e1.record()
        torch.cuda.synchronize()
        line += f"  {'brick' if brick else 'tap  '} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us"
    print(line)
ops.set_tuning("wgrad_brick", -1)
