"""A/B of ofsv_conv_tc's paired-tile mode and pipeline depth (ofsv_set_tuning("tc_pair", 0 | -1 | 1), ("tc_stages", 0 | 2..4)) on
UPFlowNet.forward_2_frame_v3, 8 pairs of 256 x 832.  Measured on a B200 (conv_tc ms per call): one tile / 4 stages 4.42; one tile,
depth for two CTAs per SM 4.15; paired where it pays + that depth (the defaults) 4.03; pairs everywhere 4.29; 2 stages 4.57."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops
from opticalflowscivis_b200.upflow.net import UPFlowNet
torch.manual_seed(0)
net = UPFlowNet().cuda()
a, b = torch.rand(8, 3, 256, 832, device='cuda') - 0.5, torch.rand(8, 3, 256, 832, device='cuda') - 0.5
for mode, stages in ((0, 4), (0, 0), (-1, 0), (-1, 4), (1, 0), (0, 4), (0, 0), (0, 2), (0, 3)):
    ops.set_tuning("tc_pair", mode)
    ops.set_tuning("tc_stages", stages)
    for _ in range(3): net.forward_2_frame_v3(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.TIMER = t = ops.LaunchTimer()
    e0.record()
    for _ in range(5): net.forward_2_frame_v3(a, b)
    e1.record(); torch.cuda.synchronize()
    ops.TIMER = None
    tot = t.totals()
    print(mode, stages, round(e0.elapsed_time(e1) / 5, 3), {k: round(v[1] / 5, 3) for k, v in tot.items() if k.startswith('conv')})
