"""Pins the warp BACKWARD oracle (torch autograd through oracle/ops_ref.warp{2,3}d_ref) against autograd through the
reference's own `warp` (Flow-2D|3D/model/warplayer.py, imported unmodified from /root/reference) and writes
tests/golden/warp_bwd.npz.  Build container only:

    python tests/golden/make_warp_bwd_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg                                              # noqa: E402
from oracle import ops_ref                                            # noqa: E402


def cases(nd):
    g = torch.Generator().manual_seed(4321 + nd)
    shapes = [(2, 3, 20, 28), (1, 1, 33, 47)] if nd == 2 else [(2, 2, 6, 8, 10), (1, 1, 12, 20, 16)]
    for shp in shapes:
        n, c, *sp = shp
        fs = (n, nd, *sp)
        src = torch.rand(shp, generator=g)
        yield "rand", src, torch.randn(fs, generator=g) * 3.0
        yield "frac", src, torch.rand(fs, generator=g) * 0.9 + 0.05          # stays inside the cell: smooth gradient
        yield "far", src, torch.randn(fs, generator=g) * 500.0               # clipped: zero flow gradient
        edge = torch.zeros(fs)
        for a in range(nd):
            edge[:, a] = float(max(sp))
        yield "edge", src, edge


def main():
    torch.manual_seed(1234)
    fix, report = {}, []
    for nd in (2, 3):
        _, warplayer = mg.load_flow(nd)
        warplayer.device = torch.device("cpu")
        ref_fn = ops_ref.warp2d_ref if nd == 2 else ops_ref.warp3d_ref
        g = torch.Generator().manual_seed(77 + nd)
        for i, (name, src, flow) in enumerate(cases(nd)):
            go = torch.randn(src.shape, generator=g)
            a, b = src.clone().requires_grad_(), flow.clone().requires_grad_()
            warplayer.warp(a, b).backward(go)
            a2, b2 = src.clone().requires_grad_(), flow.clone().requires_grad_()
            ref_fn(a2, b2).backward(go)
            assert torch.equal(a.grad, a2.grad) and torch.equal(b.grad, b2.grad), f"warp{nd}d backward restatement != reference ({name})"
            k = f"w{nd}_{i}_{name}"
            fix[f"{k}_src"], fix[f"{k}_flow"], fix[f"{k}_gout"] = src.numpy(), flow.numpy(), go.numpy()
            fix[f"{k}_gsrc"], fix[f"{k}_gflow"] = a.grad.numpy(), b.grad.numpy()
            report.append(f"warp{nd}d backward {name} {tuple(src.shape)}: autograd(oracle) bit-exact vs autograd(reference)")
    np.savez_compressed(os.path.join(HERE, "warp_bwd.npz"), **fix)
    with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
