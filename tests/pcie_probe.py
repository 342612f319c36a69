"""PCIe copy bandwidth and copy/compute overlap probe (explains bench.py's e2e vs value gap)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
for mb in (32, 128, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {mb} MiB pinned: {5 * mb / 1024 / (e0.elapsed_time(e1) / 1e3):.1f} GiB/s")
# bidirectional on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h1 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d1 = torch.empty(256 << 20, dtype=torch.uint8, device=dev); d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"bidirectional 256 MiB each way: {5 * 0.25 / dt:.1f} GiB/s per direction")
