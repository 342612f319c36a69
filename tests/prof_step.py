"""Tiny driver for ncu captures: warm-up inferences, then ONE Model.inference inside a cudaProfilerStart/Stop range
(run ncu with --profile-from-start off) on synthetic pairs.  usage: prof_step.py [size=256] [warmup=3] [2d|3d] [pairs=1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import synth  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 256
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
two_d = len(sys.argv) > 3 and sys.argv[3] == "2d"
pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
torch.manual_seed(1234)
if two_d:
    from opticalflowscivis_b200.flow2d.model.RIFE import Model
    a, _, b = synth.droplet2d(64)
    d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
else:
    from opticalflowscivis_b200.flow3d.model.RIFE import Model
    a, _, b = synth.droplet3d_u8(pairs, s)
    d0, d1 = torch.from_numpy(a).cuda().float() / 255, torch.from_numpy(b).cuda().float() / 255
m = Model()
m.eval()
for _ in range(warm):
    out = m.inference(d0, d1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m.inference(d0, d1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
