"""Small end-to-end exercise of every kernel on ragged shapes, meant to be run under `compute-sanitizer --tool memcheck`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops
from opticalflowscivis_b200.flow2d.model.RIFE import Model as M2
from opticalflowscivis_b200.flow3d.model.RIFE import Model as M3
dev = torch.device("cuda")
g = torch.Generator().manual_seed(3)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
ops.warp3d(r(1, 2, 9, 33, 14).abs(), r(1, 3, 9, 33, 14) * 3)
ops.warp2d(r(2, 3, 17, 33), r(2, 2, 17, 33) * 3)
ops.warp_blend(r(1, 1, 8, 40, 12), r(1, 1, 8, 40, 12), r(1, 6, 8, 40, 12), r(1, 1, 8, 40, 12))
f1, f2 = r(2, 7, 9, 11), r(2, 7, 9, 11)
o = ops.corr81_fwd(f1, f2, leaky_slope=0.1)
ops.corr81_bwd(f1, f2, torch.randn_like(o))
# correlation: channel split + finalize (coarse levels), the 16-byte cp.async path (W % 4 == 0) and ragged channel counts
for shp in ((3, 50, 4, 13), (2, 37, 8, 24), (1, 96, 16, 52)):
    a, b = r(*shp), r(*shp)
    o = ops.corr81_fwd(a, b, leaky_slope=0.1)
    ops.corr81_bwd(a, b, torch.randn_like(o))
# stage kernels on the H-fastest state: every (scale_head, scale_next) pair, ragged H, with and without the planar outputs
for (sh, sn, prev) in ((4, 2, False), (2, 1, True), (1, 0, True), (1, 1, True), (2, 2, True), (4, 0, False), (0, 0, True), (0, 2, True)):
    n, sp = 2, (16, 48, 40)
    i0, i1 = r(n, 1, *sp).abs(), r(n, 1, *sp).abs()
    fm_prev = (r(n, sp[0], sp[2], sp[1], 8) * 2).contiguous() if (prev or sh == 0) else None
    head = r(n, sp[0] // sh, sp[2] // sh, sp[1] // sh, 8) if sh else None
    for want in (True, False):
        ops.block_stage_3d(head, fm_prev, i0, i1, sh, sn, want, want, pack_s2d=bool(sn), key="san", hfast=True)
ops.upsample_flow_ac(r(1, 2, 5, 7), 20, 28)
ops.warping_no_div(r(1, 4, 12, 20), r(1, 2, 12, 20))
ops.u8_to_f32(torch.randint(0, 255, (1000003,), dtype=torch.uint8, device=dev)[:999985].contiguous() if False else torch.randint(0, 255, (999984 + 16,), dtype=torch.uint8, device=dev))
torch.manual_seed(0)
m3 = M3(); m3.eval()
m3.inference(torch.rand(2, 1, 16, 48, 32, device=dev), torch.rand(2, 1, 16, 48, 32, device=dev))
m3f = M3(precision="fp32"); m3f.eval()
m3f.inference(torch.rand(1, 1, 16, 16, 32, device=dev), torch.rand(1, 1, 16, 16, 32, device=dev))
m2 = M2(); m2.eval()
m2.inference(torch.rand(3, 1, 32, 48, device=dev), torch.rand(3, 1, 32, 48, device=dev))
torch.cuda.synchronize()
print("sanitize_small ok")
