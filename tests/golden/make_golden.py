"""Pins the oracle against the reference ITSELF and writes the golden fixtures in this directory.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

For every case it (1) runs the unmodified reference module imported from /root/reference with the
loader recipe of SURVEY.md Appendix C, (2) runs the restatements in oracle/ (torch and plain C) on the
same inputs, (3) asserts they agree (bit-exactly where stated), and (4) stores inputs/outputs as small
.npz fixtures that tests/test_oracle_golden.py replays without the reference.
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

from oracle import c_oracle as co                                    # noqa: E402
from oracle import ops_ref                                           # noqa: E402
from oracle.ifnet_ref import IFNetRef                                # noqa: E402


def _purge():
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]


def load_flow(nd: int):
    """stub `utils` (plotting only), import model.IFNet from Flow-2D / Flow-3D."""
    _purge()
    sys.path.insert(0, f"{REF}/Flow-{nd}D")
    stub = types.ModuleType("utils")
    for n in ("plot_loss", "visualize_ind", "visualize_series", "visualize_series_flow", "visualize_large"):
        setattr(stub, n, lambda *a, **k: None)
    sys.modules["utils"] = stub
    ifnet = importlib.import_module("model.IFNet")
    warplayer = importlib.import_module("model.warplayer")
    return ifnet, warplayer


def load_upflow():
    _purge()
    sys.path.insert(0, f"{REF}/UPFlow")
    for n in ("imageio", "png", "correlation_cuda"):
        sys.modules.setdefault(n, types.ModuleType(n))
    pwc = importlib.import_module("model.pwc_modules")
    corr = importlib.import_module("utils.pytorch_correlation")
    return pwc, corr


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def warp_cases(nd: int):
    """Adversarial warp inputs: random, zero flow, integer shifts, far out of bounds, landing on S-1."""
    g = torch.Generator().manual_seed(1234 + nd)
    shapes = [(2, 3, 20, 28), (1, 1, 160, 224)] if nd == 2 else [(2, 2, 6, 8, 10), (1, 1, 16, 16, 16), (1, 1, 24, 16, 20)]
    for shp in shapes:
        n, c, *sp = shp
        src = torch.rand(shp, generator=g)
        fshape = (n, nd, *sp)
        yield "rand", src, torch.randn(fshape, generator=g) * 3.0
        yield "zero", src, torch.zeros(fshape)
        yield "int", src, torch.randint(-4, 5, fshape, generator=g).float()
        yield "far", src, torch.randn(fshape, generator=g) * 500.0
        edge = torch.zeros(fshape)
        for a in range(nd):   # push every sample exactly onto / past the last index of the sampled axis
            edge[:, a] = float(max(sp))
        yield "edge", src, edge
    # binary {0,1} volume (droplet-like): steepest gradients, SURVEY App. A
    shp = (1, 1, 64, 96) if nd == 2 else (1, 1, 16, 24, 32)
    src = (torch.rand(shp, generator=g) > 0.5).float()
    yield "binary", src, torch.randn((1, nd, *shp[2:]), generator=g) * 2.0


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    report = []

    # ---------------- warp (a1, a2) ----------------
    for nd in (2, 3):
        _, warplayer = load_flow(nd)
        warplayer.device = torch.device("cpu")
        fix = {}
        for i, (name, src, flow) in enumerate(warp_cases(nd)):
            ref = warplayer.warp(src, flow)
            t = (ops_ref.warp2d_ref if nd == 2 else ops_ref.warp3d_ref)(src, flow)
            cc = (co.warp2d if nd == 2 else co.warp3d)(src.numpy(), flow.numpy())
            assert torch.equal(ref, t), f"warp{nd}d torch restatement != reference ({name})"
            d = np.abs(cc - ref.numpy()).max()
            assert d == 0.0, f"warp{nd}d C restatement != reference ({name}): {d}"
            if src.numel() <= 40000:
                fix[f"{i}_{name}_src"], fix[f"{i}_{name}_flow"], fix[f"{i}_{name}_out"] = src.numpy(), flow.numpy(), ref.numpy()
            report.append(f"warp{nd}d {name} {tuple(src.shape)}: torch bit-exact, C bit-exact")
        if nd == 3:   # zero flow on a cube rotates axes
            x = torch.rand(1, 1, 8, 8, 8)
            assert torch.allclose(warplayer.warp(x, torch.zeros(1, 3, 8, 8, 8)), x.permute(0, 1, 3, 4, 2), atol=1e-6)
        np.savez_compressed(os.path.join(HERE, f"warp{nd}d.npz"), **fix)

    # ---------------- IFNet / Model.inference (a3-a7) ----------------
    for nd, shp in ((2, (2, 1, 32, 64)), (3, (1, 1, 16, 16, 32))):
        ifnet, _ = load_flow(nd)
        torch.manual_seed(1234)
        ref_net = quiet(ifnet.IFNet).eval()
        torch.manual_seed(1234)
        mine = IFNetRef(nd).eval()
        sd_ref, sd = ref_net.state_dict(), mine.state_dict()
        assert list(sd_ref.keys()) == list(sd.keys()), "state_dict keys differ"
        assert all(torch.equal(sd_ref[k], sd[k]) for k in sd), "seeded init differs from the reference"
        g = torch.Generator().manual_seed(99)
        img0 = torch.rand(shp, generator=g)
        img1 = torch.roll(img0, shifts=2, dims=-1) * 0.9 + 0.05
        with torch.no_grad():
            fl_r, mk_r, mg_r, *_ = quiet(ref_net, torch.cat((img0, img1), 1), [4, 2, 1])
            fl, mk, mg = mine(torch.cat((img0, img1), 1), (4, 2, 1))
        if nd == 3:
            mk_r = [None, None, mk_r]          # 3D returns mask_list[2] only (Flow-3D/model/IFNet.py:280)
        for i in range(3):
            assert torch.equal(fl_r[i], fl[i]) and torch.equal(mg_r[i], mg[i]), f"IFNet{nd}D scale {i} differs"
            if mk_r[i] is not None:
                assert torch.equal(mk_r[i], mk[i])
        wsum = float(sum(v.double().abs().sum() for v in sd.values()))
        np.savez_compressed(os.path.join(HERE, f"ifnet{nd}d.npz"), img0=img0.numpy(), img1=img1.numpy(),
                            weight_abs_sum=np.float64(wsum), seed=np.int64(1234),
                            **{f"flow{i}": fl[i].numpy() for i in range(3)},
                            **{f"mask{i}": mk[i].numpy() for i in range(3)},
                            **{f"merged{i}": mg[i].numpy() for i in range(3)})
        report.append(f"IFNet{nd}D {shp}: restatement bit-exact vs reference (flow/mask/merged, 3 scales), |w|_1={wsum:.6f}")

    # ---------------- UPFlow operators (a8-a11) ----------------
    pwc, corr = load_upflow()
    g = torch.Generator().manual_seed(7)
    fix = {}
    cp = corr.Corr_pyTorch(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    for i, (b, c, h, w) in enumerate(((2, 32, 16, 24), (1, 196, 4, 13), (2, 7, 9, 11))):
        f1, f2 = torch.randn(b, c, h, w, generator=g), torch.randn(b, c, h, w, generator=g)
        ref = cp(f1, f2)
        t = ops_ref.corr81_ref(f1, f2)
        cc = co.corr81(f1.numpy(), f2.numpy())
        assert (ref - t).abs().max() < 2e-6 and np.abs(cc - ref.numpy()).max() < 2e-6
        # backward: autograd through the reference twin vs the C restatement
        f1g, f2g = f1.clone().requires_grad_(), f2.clone().requires_grad_()
        go = torch.randn(ref.shape, generator=g)
        cp(f1g, f2g).backward(go)
        g1, g2 = co.corr81_bwd(f1.numpy(), f2.numpy(), go.numpy())
        assert np.abs(g1 - f1g.grad.numpy()).max() < 2e-5 and np.abs(g2 - f2g.grad.numpy()).max() < 2e-5
        fix.update({f"corr{i}_f1": f1.numpy(), f"corr{i}_f2": f2.numpy(), f"corr{i}_out": ref.numpy(),
                    f"corr{i}_gout": go.numpy(), f"corr{i}_g1": f1g.grad.numpy(), f"corr{i}_g2": f2g.grad.numpy()})
        report.append(f"corr81 {(b, c, h, w)}: |ref-torch|,|ref-C| < 2e-6; bwd < 2e-5")
    for i, (b, h_, w_, h, w) in enumerate(((2, 4, 13, 8, 26), (1, 16, 52, 64, 208), (1, 5, 7, 5, 7))):
        fl = torch.randn(b, 2, h_, w_, generator=g) * 4
        tgt = torch.empty(b, 1, h, w)
        ref = pwc.upsample2d_flow_as(fl, tgt, mode="bilinear", if_rate=True)
        t = ops_ref.upsample2d_flow_as_ref(fl, h, w)
        cc = co.upsample_flow_ac(fl.numpy(), h, w)
        assert torch.equal(ref, t) and np.abs(cc - ref.numpy()).max() < 4e-6
        fix.update({f"ups{i}_in": fl.numpy(), f"ups{i}_out": ref.numpy()})
        report.append(f"upsample2d_flow_as {(b, h_, w_)}->{(h, w)}: torch bit-exact, C < 4e-6")
    wl = pwc.WarpingLayer_no_div()
    import warnings
    warnings.simplefilter("ignore")
    for i, (b, c, h, w, kind) in enumerate(((2, 3, 20, 28, "rand"), (1, 8, 16, 52, "int"), (1, 2, 8, 26, "far"))):
        x = torch.rand(b, c, h, w, generator=g)
        fl = torch.randn(b, 2, h, w, generator=g) * (50.0 if kind == "far" else 3.0)
        if kind == "int":
            fl = fl.round()
        ref = wl(x, fl.clone())
        t = ops_ref.warping_layer_no_div_ref(x, fl)
        cc = co.warping_no_div(x.numpy(), fl.numpy())
        assert torch.equal(ref, t) and np.array_equal(cc, ref.numpy()), f"WarpingLayer_no_div {kind}"
        fix.update({f"wnd{i}_x": x.numpy(), f"wnd{i}_flow": fl.numpy(), f"wnd{i}_out": ref.numpy()})
        report.append(f"WarpingLayer_no_div {kind} {(b, c, h, w)}: torch bit-exact, C bit-exact (incl. >=1 validity mask)")
    np.savez_compressed(os.path.join(HERE, "upflow_ops.npz"), **fix)

    with open(os.path.join(HERE, "PINNING.txt"), "w") as f:
        f.write("oracle pinned against /root/reference (imported unmodified) with torch %s\n" % torch.__version__)
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
