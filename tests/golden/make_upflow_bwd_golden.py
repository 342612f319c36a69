"""Golden vectors for the backward of a10 / a11: autograd through the reference's own `upsample2d_flow_as` and
`WarpingLayer_no_div` (UPFlow/model/pwc_modules.py, imported unmodified from /root/reference), checked bit-exact against
autograd through the oracle restatements (oracle/ops_ref.py).  Build container only:

    python tests/golden/make_upflow_bwd_golden.py
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg                                              # noqa: E402
from oracle import ops_ref                                            # noqa: E402


def main():
    warnings.simplefilter("ignore")
    pwc, _ = mg.load_upflow()
    g = torch.Generator().manual_seed(4242)
    fix, report = {}, []
    for i, (b, h_, w_, h, w) in enumerate(((2, 4, 13, 8, 26), (1, 16, 52, 64, 208), (1, 5, 7, 5, 7), (2, 6, 9, 17, 23))):
        fl = torch.randn(b, 2, h_, w_, generator=g) * 4
        go = torch.randn(b, 2, h, w, generator=g)
        a = fl.clone().requires_grad_()
        pwc.upsample2d_flow_as(a, torch.empty(b, 1, h, w), mode="bilinear", if_rate=True).backward(go)
        a2 = fl.clone().requires_grad_()
        ops_ref.upsample2d_flow_as_ref(a2, h, w).backward(go)
        assert torch.equal(a.grad, a2.grad), "upsample2d_flow_as backward restatement != reference"
        fix.update({f"ups{i}_in": fl.numpy(), f"ups{i}_gout": go.numpy(), f"ups{i}_gin": a.grad.numpy()})
        report.append(f"upsample2d_flow_as backward {(b, h_, w_)}->{(h, w)}: autograd(oracle) bit-exact vs autograd(reference)")
    wl = pwc.WarpingLayer_no_div()
    for i, (b, c, h, w, kind) in enumerate(((2, 3, 20, 28, "rand"), (1, 8, 16, 52, "frac"), (1, 2, 8, 26, "far"), (2, 4, 13, 17, "rand"))):
        x = torch.rand(b, c, h, w, generator=g)
        fl = torch.randn(b, 2, h, w, generator=g) * (50.0 if kind == "far" else 3.0)
        if kind == "frac":
            fl = torch.rand(b, 2, h, w, generator=g) * 0.9 + 0.05
        go = torch.randn(b, c, h, w, generator=g)
        a, f = x.clone().requires_grad_(), fl.clone().requires_grad_()
        wl(a, f).backward(go)
        a2, f2 = x.clone().requires_grad_(), fl.clone().requires_grad_()
        ops_ref.warping_layer_no_div_ref(a2, f2).backward(go)
        assert torch.equal(a.grad, a2.grad) and torch.equal(f.grad, f2.grad), f"WarpingLayer_no_div backward restatement != reference ({kind})"
        fix.update({f"wnd{i}_x": x.numpy(), f"wnd{i}_flow": fl.numpy(), f"wnd{i}_gout": go.numpy(),
                    f"wnd{i}_gx": a.grad.numpy(), f"wnd{i}_gflow": f.grad.numpy()})
        report.append(f"WarpingLayer_no_div backward {kind} {(b, c, h, w)}: autograd(oracle) bit-exact vs autograd(reference)")
    np.savez_compressed(os.path.join(HERE, "upflow_bwd.npz"), **fix)
    with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
