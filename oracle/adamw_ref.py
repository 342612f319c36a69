"""numpy float32 restatement of the optimizer step of the reference's training loop.  TEST INFRASTRUCTURE (see
oracle/__init__.py).

The reference does not implement the optimizer itself: it calls `torch.optim.AdamW(self.flownet.parameters(), lr=1e-6,
weight_decay=1e-3)` (Flow-2D/model/RIFE.py:26, Flow-3D/model/RIFE.py:29), sets `param_group['lr']` every step (:81-82 /
:86-87) and calls `optimG.step()` (:317 / :259).  The arithmetic restated here is torch 2.x `_single_tensor_adamw`
(amsgrad=False, maximize=False); tests/golden/make_adamw_golden.py pins it against torch.optim.AdamW itself on CPU.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-3):
    """One AdamW step on float32 arrays (updated copies are returned): `step` is the count AFTER this update (>= 1)."""
    p, g, m, v = (np.asarray(a, dtype=f32) for a in (p, g, m, v))
    p = (p * f32(1.0 - lr * weight_decay)).astype(f32)                 # param.mul_(1 - lr * weight_decay)
    m = (m + (g - m) * f32(1.0 - beta1)).astype(f32)                   # exp_avg.lerp_(grad, 1 - beta1)
    v = (v * f32(beta2) + f32(1.0 - beta2) * g * g).astype(f32)        # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / f32(math.sqrt(bc2)) + f32(eps)).astype(f32)   # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = (p - f32(step_size) * (m / denom)).astype(f32)                 # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v
