"""Tiny driver for ncu captures: a few Model.inference calls on one synthetic 256^3 (or given size) pair."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import synth  # noqa: E402
from opticalflowscivis_b200.flow3d.model.RIFE import Model  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(1234)
m = Model()
m.eval()
a, _, b = synth.droplet3d_u8(1, s)
d0, d1 = torch.from_numpy(a).cuda().float() / 255, torch.from_numpy(b).cuda().float() / 255
for _ in range(iters):
    out = m.inference(d0, d1)
torch.cuda.synchronize()
print("ok", float(out[0].mean()))
