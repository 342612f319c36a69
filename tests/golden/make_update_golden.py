"""Pins oracle/train_ref.py (Model.update, the training step) against the reference ITSELF and writes tests/golden/update.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_update_golden.py

For nd = 3 and nd = 2: construct the unmodified reference `Model` on CPU under seed 1234 (SURVEY.md Appendix C recipe), copy its
weights into the oracle's IFNetRef, run THREE `update(..., learning_rate=1e-4, training=True)` steps on the same seeded triplet with
both, assert that losses and parameters agree, and store the losses, per-tensor first-step gradient norms and per-tensor parameter
deltas so that tests/test_oracle_golden.py can replay the oracle without the reference.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.ifnet_ref import IFNetRef                    # noqa: E402
from oracle.train_ref import TrainerRef, training_triplet  # noqa: E402

CASES = {3: dict(n=2, size=32), 2: dict(n=2, size=64)}
LR, STEPS = 1e-4, 3


def _purge():
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]


def load_rife(nd):
    _purge()
    sys.path.insert(0, f"{REF}/Flow-{nd}D")
    stub = types.ModuleType("utils")
    for n in ("plot_loss", "visualize_ind", "visualize_series", "visualize_series_flow", "visualize_large"):
        setattr(stub, n, lambda *a, **k: None)
    sys.modules["utils"] = stub
    with contextlib.redirect_stdout(io.StringIO()):
        rife = importlib.import_module("model.RIFE")
    rife.device = torch.device("cpu")
    return rife


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    torch.set_num_threads(os.cpu_count())
    out, log = {}, []
    for nd, cfg in CASES.items():
        rife = load_rife(nd)
        torch.manual_seed(1234)
        ref = quiet(rife.Model)
        torch.manual_seed(1234)
        net = IFNetRef(nd)
        sd_ref = ref.flownet.state_dict()
        same_init = all(torch.equal(v, sd_ref[k]) for k, v in net.state_dict().items())
        net.load_state_dict(sd_ref)
        orc = TrainerRef(nd, net)
        img0, img1, gt = training_triplet(nd, cfg["n"], cfg["size"])
        imgs = torch.cat((img0, img1), 1)
        p0 = {k: v.clone() for k, v in sd_ref.items()}
        names = [k for k, _ in net.named_parameters()]
        for step in range(STEPS):
            if nd == 3:
                _, ir = quiet(ref.update, imgs, gt, learning_rate=LR, training=True)
            else:
                _, ir = quiet(ref.update, imgs, gt, "droplet2d", learning_rate=LR, training=True)
            _, io_ = orc.update(imgs, gt, learning_rate=LR, training=True)
            for key in ("loss_l1", "loss_tea", "loss_distill", "loss_G"):
                a, b = float(ir[key]), float(io_[key])
                assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (nd, step, key, a, b)
                out[f"nd{nd}_step{step}_{key}"] = np.float64(b)
            if step == 0:
                gr = dict(ref.flownet.named_parameters())
                for k, p in net.named_parameters():
                    d = (p.grad - gr[k].grad).abs().max().item()
                    assert d <= 1e-6 * max(1e-3, gr[k].grad.abs().max().item()) + 1e-9, (nd, k, d)
                out[f"nd{nd}_gradnorm"] = np.array([p.grad.double().norm().item() for _, p in net.named_parameters()])
                out[f"nd{nd}_gradhead"] = np.stack([p.grad.flatten()[:8].double().numpy() if p.numel() >= 8 else
                                                    np.pad(p.grad.flatten().double().numpy(), (0, 8 - p.numel())) for _, p in net.named_parameters()])
        sd_o, sd_r = net.state_dict(), ref.flownet.state_dict()
        worst = max((sd_o[k] - sd_r[k]).abs().max().item() for k in sd_o)
        assert worst <= 2e-7, (nd, worst)
        out[f"nd{nd}_deltanorm"] = np.array([(sd_o[k] - p0[k]).double().norm().item() for k in names])
        out[f"nd{nd}_deltasum"] = np.array([(sd_o[k] - p0[k]).double().sum().item() for k in names])
        out[f"nd{nd}_w1"] = np.float64(sum(v.double().abs().sum().item() for v in p0.values()))
        log.append(f"Model.update nd={nd} N={cfg['n']} size={cfg['size']}: 3 steps, losses |ref-oracle| <= 1e-6 rel, step-0 gradients <= 1e-6 rel, "
                   f"parameters after 3 steps max |ref-oracle| = {worst:.1e}; seeded init identical to the reference's: {same_init}; "
                   f"loss_G = {[round(float(out[f'nd{nd}_step{s}_loss_G']), 6) for s in range(STEPS)]}")
        print(log[-1])
    out["names3"] = np.array([k for k, _ in IFNetRef(3).named_parameters()])
    np.savez_compressed(os.path.join(HERE, "update.npz"), **out)
    with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
        for line in log:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
