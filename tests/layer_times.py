"""Per-launch CUDA-event times of one Model.inference (median over repetitions), in launch order.
usage: layer_times.py [size=256] [pairs=1] [reps=5]      (set OFSV_LIB=<other build> to A/B two builds of libofsv)"""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import ops, synth
from opticalflowscivis_b200.flow3d.model.RIFE import Model
s = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(1234)
m = Model(); m.eval()
a, _, b = synth.droplet3d_u8(n, s)
d0, d1 = torch.from_numpy(a).cuda().float() / 255, torch.from_numpy(b).cuda().float() / 255
for _ in range(3):
    m.inference(d0, d1)
torch.cuda.synchronize()
runs = []
for _ in range(reps):
    t = ops.LaunchTimer(); ops.TIMER = t
    m.inference(d0, d1)
    torch.cuda.synchronize(); ops.TIMER = None
    runs.append([(nm, x.elapsed_time(y) * 1e3) for nm, x, y in t.seq])
names = [nm for nm, _ in runs[0]]
med = [statistics.median(r[i][1] for r in runs) for i in range(len(names))]
tag = os.environ.get("OFSV_LIB", "default")
print(f"== {tag}: {s}^3 x{n}, total {sum(med):.0f} us")
print(" ".join(f"{nm.replace('conv_','c')[:6]}:{v:.0f}" for nm, v in zip(names, med)))
