"""Golden record of the reference's UPFlow network for SURVEY.md §8 f.2 — writes tests/golden/upflow_net.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_upflow_net_golden.py
Imports the unmodified UPFlow/model/upflow.py on CPU (stub imageio / png / correlation_cuda, device = cpu,
`if_use_cor_pytorch=True`), loads oracle.upflow_ref.deterministic_state weights, runs `forward_2_frame_v3` on a seeded pair with
the training configuration of scripts/simple_train.py:321-329 (per-plane feature normalisation, no sgu) and with test.py:116-121's
(sgu up-sampling on), and records the flow pyramids and the sub-sampled full-resolution flows.  Also pins the operator
restatements of oracle/upflow_ref.py (normalize_features, torch_warp, occlusion check) against the reference functions.
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

from oracle import upflow_ref as ur          # noqa: E402

SHAPE = (1, 128, 192)


def load():
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]
    sys.path.insert(0, f"{REF}/UPFlow")
    for n in ("imageio", "png", "correlation_cuda", "cv2"):
        sys.modules.setdefault(n, types.ModuleType(n))
    with contextlib.redirect_stdout(io.StringIO()):
        up = importlib.import_module("model.upflow")
    up.device = torch.device("cpu")
    return up


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    up = load()
    tools = importlib.import_module("utils.tools").tools
    out, log = {}, []
    # operator restatements
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn((2, 5, 6, 9), generator=g), torch.randn((2, 5, 6, 9), generator=g) * 3 + 1
    ra = up.network_tools.normalize_features((a, b), normalize=True, center=True, moments_across_channels=False, moments_across_images=False)
    oa = ur.normalize_features_ref(a, b)
    assert all(torch.equal(x, y) for x, y in zip(ra, oa))
    x, fl = torch.randn((2, 3, 12, 20), generator=g), torch.randn((2, 2, 12, 20), generator=g) * 3
    assert torch.equal(quiet(tools.torch_warp, x, fl), ur.torch_warp_ref(x, fl))
    f1, f2 = torch.randn((2, 2, 12, 20), generator=g) * 2, torch.randn((2, 2, 12, 20), generator=g) * 2
    occ = tools.occ_check_model(occ_type="for_back_check", occ_alpha_1=0.1, occ_alpha_2=0.5, obj_out_all="obj")
    rf, rb = quiet(occ, flow_f=f1, flow_b=f2)
    of, ob = ur.occ_check_ref(f1, f2)
    assert torch.equal(rf, of) and torch.equal(rb, ob)
    log.append("UPFlow operators: normalize_features (per-plane moments), tools.torch_warp, forward-backward occlusion check: restatements bit-exact vs reference")
    b_, h, w = SHAPE
    im1, im2 = ur.smooth_pair(b_, h, w)
    for tag, sgu in (("train", False), ("sgu", True)):
        conf = up.UPFlow_net.config()
        conf.update({"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
                     "if_sgu_upsample": sgu, "if_use_cor_pytorch": True}) if hasattr(conf, "update") else None
        for k, v in (("if_norm_before_cost_volume", True), ("norm_moments_across_channels", False), ("norm_moments_across_images", False),
                     ("if_sgu_upsample", sgu), ("if_use_cor_pytorch", True)):
            setattr(conf, k, v)
        torch.manual_seed(0)
        net = quiet(up.UPFlow_net, conf)
        net.eval()
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(ur.deterministic_state(shapes))
        with torch.no_grad():
            ff, fb, flows = quiet(net.forward_2_frame_v3, im1, im2)
        out[f"{tag}_names"] = np.array(sorted(shapes))
        out[f"{tag}_shapes"] = np.array([str(shapes[k]) for k in sorted(shapes)])
        out[f"{tag}_flow_f_sub4"] = ff[:, :, ::4, ::4].numpy()
        out[f"{tag}_flow_b_sub4"] = fb[:, :, ::4, ::4].numpy()
        for i, (a_, b__) in enumerate(flows):
            out[f"{tag}_lvl{i}_f"], out[f"{tag}_lvl{i}_b"] = a_.numpy(), b__.numpy()
        of, ob = quiet(occ, flow_f=ff, flow_b=fb)
        out[f"{tag}_occ_f_mean"], out[f"{tag}_occ_b_mean"] = np.float64(of.mean()), np.float64(ob.mean())
        log.append(f"UPFlow_net.forward_2_frame_v3 ({tag}: sgu={sgu}) {SHAPE}: {len(shapes)} tensors / {sum(int(np.prod(s)) for s in shapes.values())} parameters, "
                   f"mean |flow_f| = {float(ff.abs().mean()):.3f} px, recorded 5 pyramid levels + the full-resolution flows (every 4th pixel)")
    np.savez_compressed(os.path.join(HERE, "upflow_net.npz"), **out)
    with open(os.path.join(HERE, "PINNING.txt"), "a") as f:
        for line in log:
            print(line)
            f.write(line + "\n")


if __name__ == "__main__":
    main()
