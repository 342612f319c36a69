"""compute-sanitizer driver for the next-tier kernels: one small `Model.update` step in 3-D and 2-D (eager), one UPFlowNet forward with
the sgu model, the refinement nets.  usage: [compute-sanitizer --tool memcheck] python tests/sanitize_train.py  (all three weight-gradient kernel policies, ragged shapes)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops, refine  # noqa: E402
from opticalflowscivis_b200.rife import Model2D, Model3D  # noqa: E402
from opticalflowscivis_b200.upflow.net import UPFlowNet, occ_check  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for nd, sp in ((3, (32, 32, 48)), (2, (48, 80))):
    m = (Model3D if nd == 3 else Model2D)(local_rank=-1)
    imgs, gt = torch.rand((1, 2) + sp, device=dev), torch.rand((1, 1) + sp, device=dev)
    for brick in (-1, 1, 0):
        ops.set_tuning("wgrad_brick", brick)
        _, info = m.update(imgs, gt, learning_rate=1e-5, training=True) if nd == 3 else m.update(imgs, gt, "droplet2d", learning_rate=1e-5, training=True)
    ops.set_tuning("wgrad_brick", -1)
    torch.cuda.synchronize()
    print(f"update nd={nd}: loss_G {float(info['loss_G']):.5f}")
net = UPFlowNet(if_sgu_upsample=True).to(dev)
ff, fb, flows = net.forward_2_frame_v3(torch.rand((1, 3, 64, 128), device=dev) - 0.5, torch.rand((1, 3, 64, 128), device=dev) - 0.5)
occ_check(ff, fb)
torch.cuda.synchronize()
print("upflow ok", float(ff.abs().mean()))
c, u = refine.Contextnet(2).to(dev), refine.Unet(2).to(dev)
x, fl = torch.rand((1, 1, 32, 48), device=dev), torch.randn((1, 2, 32, 48), device=dev)
c0, c1 = c(x, fl), c(x.flip(0), fl)
p = torch.rand((1, 9, 32, 48), device=dev)
u(p[:, :1], p[:, 1:2], p[:, 2:3], p[:, 3:4], p[:, 4:5], p[:, 5:], c0, c1)
torch.cuda.synchronize()
print("refine ok")
