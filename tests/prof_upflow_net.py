"""Tiny driver for ncu captures of the UPFlow network: warm-up forwards, then ONE `forward_2_frame_v3` inside a
cudaProfilerStart/Stop range (run ncu with --profile-from-start off).  usage: prof_upflow_net.py [pairs=8] [warmup=2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200.upflow.net import UPFlowNet  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(1234)
net = UPFlowNet().cuda()
g = torch.Generator().manual_seed(1234)
base = torch.nn.functional.avg_pool2d(torch.rand((b, 3, 256 + 16, 832 + 16), generator=g), 7, 1, 3)
im1, im2 = base[:, :, 8:-8, 8:-8].contiguous().cuda(), base[:, :, 8:-8, 2:-14].contiguous().cuda()
for _ in range(warm):
    net.forward_2_frame_v3(im1, im2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.forward_2_frame_v3(im1, im2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
