"""Tiny driver for ncu captures of the training tier: warm-up `Model.update` steps (eager), then ONE step inside a
cudaProfilerStart/Stop range (run ncu with --profile-from-start off).  usage: prof_train_step.py [size=64] [triplets=8] [warmup=2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200.rife import Model3D  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(1234)
m = Model3D(local_rank=-1)
g = torch.Generator().manual_seed(1234)
base = torch.nn.functional.avg_pool3d(torch.rand((n, 1, s + 8, s + 8, s + 8), generator=g), 5, 1, 2)
img0, gt, img1 = (base[:, :, 4:-4, 4:-4, o:o + s].contiguous().cuda() for o in (2, 4, 6))
imgs = torch.cat((img0, img1), 1)
for _ in range(warm):
    m.update(imgs, gt, learning_rate=3e-6, training=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.update(imgs, gt, learning_rate=3e-6, training=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
