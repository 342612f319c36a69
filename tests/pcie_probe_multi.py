"""Host-link probe for the N-GPU end-to-end figure: pinned-memory H2D / D2H / both-direction copy rates of ONE GPU
while the other ranks idle, then of ALL ranks together.  Launch with torchrun (one rank per GPU); rank 0 prints one JSON line.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/pcie_probe_multi.py"""
import json, os, sys
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "ERROR")
    dist.init_process_group("nccl", device_id=dev)
MB = 256
h_in, h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(), torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(MB << 20, dtype=torch.uint8, device=dev), torch.empty(MB << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(mode, active, reps=8):
    """GB/s per direction on this rank (0 when inactive)."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    gbs = reps * (MB << 20) / (e0.elapsed_time(e1) / 1e3) / 1e9 if active else 0.0
    barrier()
    return gbs


out = {}
for mode in ("h2d", "d2h", "both"):
    run(mode, True, 2)                                     # warm-up
    alone = run(mode, rank == 0)
    together = run(mode, True)
    t = torch.tensor([alone, together, together], device=dev, dtype=torch.float64)
    if world > 1:
        mn = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        t[2] = mn[2]
    out[mode] = {"one_gpu_alone_GBps_per_direction": round(float(t[0]), 1), "all_gpus_sum_GBps_per_direction": round(float(t[1]), 1),
                 "slowest_rank_GBps_per_direction": round(float(t[2]), 1)}
if rank == 0:
    print(json.dumps({"probe": "pinned host <-> device copies, 256 MiB blocks", "n_gpus": world, "host_cpus": os.cpu_count(), **out}), flush=True)
if world > 1:
    dist.destroy_process_group()
