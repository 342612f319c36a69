"""Is Model.inference host/launch-bound?  Eager per-call time vs replay of a CUDA graph captured around the same call.
usage: graph_probe.py  (prints one line per workload)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opticalflowscivis_b200 import synth
from opticalflowscivis_b200.rife import Model2D, Model3D

dev = torch.device("cuda", 0)
for name, nd, sp, pairs in (("flow2d 64x160x224", 2, (160, 224), 64), ("flow2d 1x160x224", 2, (160, 224), 1),
                            ("flow3d 4x128^3", 3, (128,) * 3, 4), ("flow3d 1x256^3", 3, (256,) * 3, 1), ("flow3d 4x256^3", 3, (256,) * 3, 4)):
    torch.manual_seed(1234)
    model = (Model3D if nd == 3 else Model2D)(local_rank=0, precision="bf16", engine="auto")
    model.eval()
    if nd == 3:
        a, _, b = synth.droplet3d_u8(pairs, sp[0], seed=1234)
        d0, d1 = torch.from_numpy(a).to(dev).float().div_(255.0), torch.from_numpy(b).to(dev).float().div_(255.0)
    else:
        a, _, b = synth.droplet2d(pairs, *sp, seed=1234)
        d0, d1 = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    for _ in range(3):
        ref = model.inference(d0, d1)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        model.inference(d0, d1)
    e1.record()
    t_host = (time.perf_counter() - t0) / reps * 1e3          # host time to ENQUEUE one call
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            model.inference(d0, d1)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = model.inference(d0, d1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / reps
    m_ref = ref[0][2] if nd == 2 else ref[0]
    m_out = out[0][2] if nd == 2 else out[0]
    same = bool(torch.equal(m_ref, m_out))
    print(f"{name:20s}: eager {eager:7.3f} ms/call (host enqueue {t_host:6.3f} ms), graph replay {graph:7.3f} ms/call, x{eager / graph:4.2f}, identical={same}")
    del model, d0, d1, ref, out, g
    from opticalflowscivis_b200 import ops
    ops.clear_workspaces()
    torch.cuda.empty_cache()
