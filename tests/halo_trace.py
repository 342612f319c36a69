"""Debug: run block2's conv layers of one 256^3 inference with OFSV_HALO_TRACE set (timeline of CTA 0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import synth
from opticalflowscivis_b200.flow3d.model.RIFE import Model
torch.manual_seed(1234)
m = Model(); m.eval()
a, _, b = synth.droplet3d_u8(1, 256)
d0, d1 = torch.from_numpy(a).cuda().float() / 255, torch.from_numpy(b).cuda().float() / 255
m.inference(d0, d1); torch.cuda.synchronize()
os.environ["OFSV_HALO_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/halo_trace.txt"
m.inference(d0, d1); torch.cuda.synchronize()
