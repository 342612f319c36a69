"""Training-tier host pieces: the optimizer and the gradient exchange of `Model.update`.

* `FusedAdamW` — drop-in for the reference's `torch.optim.AdamW(self.flownet.parameters(), lr=1e-6, weight_decay=1e-3)`
  (Flow-2D/model/RIFE.py:26, Flow-3D/model/RIFE.py:29): same `param_groups[i]['lr']` surface (the reference sets the lr
  every step, RIFE.py:81-82 / :86-87), same arithmetic, ONE kernel launch for all parameter tensors (ofsv_adamw_step_f32).
* `allreduce_gradients` — the only collective of the path: the gradients of all parameters travel as ONE flat
  bucket through one `all_reduce` (NCCL over NVLink on GPUs, gloo in the CPU tests), like the reference's DDP wrapper
  (RIFE.py:31-32) but without per-bucket hooks; the division by the world size is folded into the optimizer (`grad_scale`).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _C

CHUNK = 4096


class FusedAdamW:
    def __init__(self, params, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3, bucket=None):
        """`bucket`: the GradientBucket holding these parameters' gradients, if any — the per-step validation of 160 gradient
        tensors (≈ 80 µs of Python, twice the kernel) then reduces to checking that the bucket views are still in place."""
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamW: no parameters")
        for p in self.params:
            if not p.is_cuda:
                raise TypeError("FusedAdamW: parameters must live on a CUDA device (no CPU path)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdamW: parameters must be contiguous float32")
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.device = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(total, device=self.device)
        self.exp_avg_sq = torch.zeros(total, device=self.device)
        self.step_count = 0
        self._table = None
        self._table_key = None
        self._bucket = bucket

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _tables(self):
        """Device tables {p, g, m, v, n} per tensor + (tensor, chunk) pairs; rebuilt when a gradient buffer moves."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in self.params)
        if key != self._table_key:
            rec = np.zeros((len(self.params), 5), dtype=np.int64)
            chunks = []
            off = 0
            for i, p in enumerate(self.params):
                n = p.numel()
                rec[i] = (p.data_ptr(), p.grad.data_ptr(), self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, n)
                chunks.extend((i, c) for c in range((n + CHUNK - 1) // CHUNK))
                off += n
            t = torch.from_numpy(rec).to(self.device)
            c = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(self.device)
            self._table, self._table_key = (t, c), key
        return self._table

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        fast = self._bucket is not None and self._table is not None and self._bucket.intact()
        for p in (() if fast else self.params):
            if p.grad is None:
                raise RuntimeError("FusedAdamW.step: every parameter needs a gradient (the reference's IFNet always produces one)")
            if not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                raise TypeError("FusedAdamW: gradients must be contiguous float32")
        g = self.param_groups[0]
        t, c = self._table if fast else self._tables()
        self.step_count += 1
        with torch.cuda.device(self.device):
            _C.check(_C.lib().ofsv_adamw_step_f32(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(c.data_ptr()), t.shape[0], c.shape[0],
                                                  float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                                  float(g["weight_decay"]), self.step_count, float(grad_scale),
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        # the kernel writes the parameters through raw pointers: tell torch, so that everything keyed on (data_ptr, _version)
        # — IFBlock's packed tap-form weights, Model's captured CUDA graphs — sees a new parameter version
        _bump_versions(self.params)


def _bump_versions(params):
    """Increment `p._version` of every parameter without touching its values (no kernel launch: an in-place no-op on a
    zero-element view shares the version counter of its base)."""
    with torch.no_grad():
        for p in params:
            p.view(-1)[:0].zero_()


class GradientBucket:
    """One flat buffer holding views for the gradients of `params`: `p.grad` of every parameter is pointed into it, so the
    whole model's gradient is one contiguous tensor for the collective (36.4 MB for the 3-D IFNet)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def intact(self) -> bool:
        """True while every `p.grad` is still the view handed out at construction (zero_grad(set_to_none=True) or an
        optimizer that re-assigns .grad would break that).  Cheap: first / last parameter only."""
        a, b = self.params[0], self.params[-1]
        return (a.grad is not None and b.grad is not None and a.grad.data_ptr() == self.flat.data_ptr()
                and b.grad.data_ptr() == self.flat.data_ptr() + 4 * (self.flat.numel() - b.numel()))


def allreduce_gradients(bucket: GradientBucket, group=None) -> float:
    """SUM-allreduce the flat gradient bucket across ranks (one collective per step); returns the `grad_scale` = 1/world to
    hand to `FusedAdamW.step` (DDP averages).  Without an initialised process group it is a no-op returning 1."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)
