"""Golden vectors for the optimizer step of the training loop: torch.optim.AdamW exactly as the reference configures and
drives it (Flow-3D/model/RIFE.py:29,86-87,259: lr set every step, weight_decay 1e-3), three steps on seeded tensors, CPU.
Checks the numpy restatement oracle/adamw_ref.py against it.

    python tests/golden/make_adamw_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.adamw_ref import adamw_step                                # noqa: E402


def main():
    g = torch.Generator().manual_seed(2024)
    shapes = [(64, 16, 3, 3, 3), (64,), (5000,), (3, 7), (1,)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in shapes]
    opt = torch.optim.AdamW(params, lr=1e-6, weight_decay=1e-3)          # RIFE.py:29
    lrs = [3e-4, 2.5e-4, 1e-6]
    fix = {f"p0_{i}": p.detach().numpy().copy() for i, p in enumerate(params)}
    mine = [(p.detach().numpy().copy(), np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)) for p in params]
    worst = 0.0
    for t, lr in enumerate(lrs, start=1):
        for pg in opt.param_groups:                                      # RIFE.py:86-87
            pg["lr"] = lr
        opt.zero_grad()
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=g) * (10.0 if i == 2 else 1.0)
            fix[f"g{t}_{i}"] = p.grad.numpy().copy()
        opt.step()                                                       # RIFE.py:259
        for i, p in enumerate(params):
            fix[f"p{t}_{i}"] = p.detach().numpy().copy()
            mine[i] = adamw_step(mine[i][0], fix[f"g{t}_{i}"], mine[i][1], mine[i][2], t, lr)
            d = np.abs(mine[i][0] - fix[f"p{t}_{i}"]).max() / max(1e-30, np.abs(fix[f"p{t}_{i}"]).max())
            worst = max(worst, float(d))
    fix["lrs"] = np.array(lrs, np.float64)
    assert worst < 1e-6, worst
    np.savez_compressed(os.path.join(HERE, "adamw.npz"), **fix)
    line = f"AdamW (torch.optim.AdamW as driven by RIFE.py:29,86-87,259), 3 steps x {len(shapes)} tensors: numpy restatement within {worst:.1e} relative"
    with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
        f.write(line + "\n")
    print(line)


if __name__ == "__main__":
    main()
