"""Can a tensor-bound halo conv and an L1-bound stage kernel share the SMs?  Times 20 convblock launches (stream 1) and
3 final-stage launches (stream 2) back to back and concurrently.  usage: [OFSV_LIB=..] [OFSV_HALO_TD=2] overlap_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import _C, ops, ifnet
torch.manual_seed(0)
dev = torch.device("cuda")
blk = ifnet.IFBlock(3, 11, 64).to(dev)
lay = blk.layers()[2]
n = 2
x = (torch.randn(n, 64, 64, 64, 64, device=dev) * 0.5).to(torch.bfloat16)
y = torch.empty_like(x)
d, _ = lay.desc(n, (64, 64, 64), _C.BF16)
fm = torch.randn(n, 256, 256, 256, 8, device=dev)
img0, img1 = torch.rand(n, 1, 256, 256, 256, device=dev), torch.rand(n, 1, 256, 256, 256, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def convs(k=20):
    for _ in range(k): ops.conv(d, x, lay.w_halo, lay.bias, lay.prelu, None, y, "halo")
def stages(k=3):
    for _ in range(k): ops.block_stage_3d(None, fm, img0, img1, 0, 0, True, True)
def timed(fn):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
convs(3); stages(1)
tc, ts = timed(convs), timed(stages)
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): convs()
    with torch.cuda.stream(s2): stages()
    cur.wait_stream(s1); cur.wait_stream(s2)
tb = timed(both)
print(f"lib={os.environ.get('OFSV_LIB','default')} td={os.environ.get('OFSV_HALO_TD','auto')}: convs {tc:.2f} ms, stages {ts:.2f} ms, serial {tc+ts:.2f} ms, concurrent {tb:.2f} ms")
