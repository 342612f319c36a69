"""Batch sharding of ensemble members / time-step pairs over the GPUs of one box.

The hot path is per-sample (no BatchNorm, no cross-sample statistics), so inference shards by contiguous ranges of pairs
with replicated weights and NO data-path collective.  The only communication is the optional gather of per-rank results /
timings at the end, done with `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n_items` owned by `rank`; the first `n_items % world` ranks get one extra item.
    Every item is owned by exactly one rank; ranks beyond `n_items` get an empty range."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"shard_range: bad arguments n_items={n_items} rank={rank} world={world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_counts(local_count: int, max_ms: float):
    """All ranks -> (total items processed, max-over-ranks device time in ms).  Uses the default process group if one is
    initialised (bench.py's reduction), otherwise returns the local values."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_count, max_ms
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    c = torch.tensor([float(local_count)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(max_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(float(c[0]))), float(t[0])


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process (one per GPU) to the CPU cores NVML reports as local to `device_index`, BEFORE any pinned host buffer
    is allocated: pinned memory is placed by first touch, and on a two-socket 8-GPU box the H2D / D2H streams of the four
    GPUs behind the other socket otherwise cross the inter-socket link (8 ranks x 400 MB per step: end-to-end throughput at 8
    GPUs was 49 % of the resident-input figure without it).  Best effort: returns what it did, never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        uuid = getattr(props, "uuid", None)
        h = None
        if uuid is not None:
            for cand in (f"GPU-{uuid}", str(uuid)):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                    break
                except Exception:  # noqa: BLE001
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        target = sorted(local & allowed)
        info.update(local_cpus=len(local), allowed_cpus=len(allowed))
        if target and len(target) < len(allowed):
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target), first=target[0], last=target[-1])
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:200]
    return info
