"""CPU: the oracle (torch restatement + plain-C restatement) replays the golden fixtures that
tests/golden/make_golden.py generated from the imported reference."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import ops_ref
from oracle.ifnet_ref import IFNetRef, ModelRef

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases(npz, suffixes):
    keys = sorted({k.rsplit("_", 1)[0] for k in npz.files})
    for k in keys:
        if all(f"{k}_{s}" in npz.files for s in suffixes):
            yield k, [npz[f"{k}_{s}"] for s in suffixes]


@pytest.mark.parametrize("nd", [2, 3])
def test_warp_golden(nd):
    z = np.load(os.path.join(G, f"warp{nd}d.npz"))
    n = 0
    for name, (src, flow, out) in _cases(z, ("src", "flow", "out")):
        t = (ops_ref.warp2d_ref if nd == 2 else ops_ref.warp3d_ref)(torch.from_numpy(src), torch.from_numpy(flow)).numpy()
        c = (co.warp2d if nd == 2 else co.warp3d)(src, flow)
        assert np.abs(t - out).max() <= 1e-6, name          # same torch build -> normally bit-exact
        assert np.abs(c - out).max() <= 1e-6, name
        # reciprocal-multiply flavour (CUDA eager) stays within ~1e-4 of the true-division reference, never equal by construction
        c2 = (co.warp2d if nd == 2 else co.warp3d)(src, flow, co.DIV_RCP)
        assert np.abs(c2 - out).max() <= 2e-4, name
        n += 1
    assert n >= 10


def test_warp3d_zero_flow_rotates_axes():
    """zero flow on a cube returns x.permute(0,1,3,4,2), not x."""
    x = np.random.default_rng(0).random((1, 1, 8, 8, 8), dtype=np.float32)
    out = co.warp3d(x, np.zeros((1, 3, 8, 8, 8), np.float32))
    assert np.abs(out - x.transpose(0, 1, 3, 4, 2)).max() < 1e-6
    assert np.abs(out - x).max() > 0.5


@pytest.mark.parametrize("nd", [2, 3])
def test_ifnet_golden(nd):
    z = np.load(os.path.join(G, f"ifnet{nd}d.npz"))
    torch.manual_seed(int(z["seed"]))
    m = ModelRef(nd).eval()
    wsum = float(sum(v.double().abs().sum() for v in m.flownet.state_dict().values()))
    assert abs(wsum - float(z["weight_abs_sum"])) < 1e-6 * wsum, "seeded init drifted (different torch build?)"
    img0, img1 = torch.from_numpy(z["img0"]), torch.from_numpy(z["img1"])
    with torch.no_grad():
        fl, mk, mg = m.flownet(torch.cat((img0, img1), 1), (4, 2, 1))
    for i in range(3):
        # conv kernels may pick a different ISA path on another host: allow 1e-4
        assert np.abs(fl[i].numpy() - z[f"flow{i}"]).max() < 1e-4
        assert np.abs(mk[i].numpy() - z[f"mask{i}"]).max() < 1e-4
        assert np.abs(mg[i].numpy() - z[f"merged{i}"]).max() < 1e-4
    out = m.inference(img0, img1)
    if nd == 3:
        assert out[0].shape == img0.shape and len(out[1]) == 3 and out[2].shape == img0.shape
    else:
        assert len(out[0]) == 3 and len(out[1]) == 3 and len(out[2]) == 3


def test_upflow_ops_golden():
    warnings.simplefilter("ignore")
    z = np.load(os.path.join(G, "upflow_ops.npz"))
    for i in range(3):
        f1, f2, out = z[f"corr{i}_f1"], z[f"corr{i}_f2"], z[f"corr{i}_out"]
        assert np.abs(ops_ref.corr81_ref(torch.from_numpy(f1), torch.from_numpy(f2)).numpy() - out).max() < 2e-6
        assert np.abs(co.corr81(f1, f2) - out).max() < 2e-6
        g1, g2 = co.corr81_bwd(f1, f2, z[f"corr{i}_gout"])
        assert np.abs(g1 - z[f"corr{i}_g1"]).max() < 2e-5 and np.abs(g2 - z[f"corr{i}_g2"]).max() < 2e-5
        lk = co.corr81(f1, f2, leaky_slope=0.1)
        assert np.allclose(lk, np.where(out > 0, out, 0.1 * out), atol=2e-6)
        fin, fout = z[f"ups{i}_in"], z[f"ups{i}_out"]
        h, w = fout.shape[2:]
        assert np.abs(ops_ref.upsample2d_flow_as_ref(torch.from_numpy(fin), h, w).numpy() - fout).max() < 1e-6
        assert np.abs(co.upsample_flow_ac(fin, h, w) - fout).max() < 4e-6
        x, fl, wo = z[f"wnd{i}_x"], z[f"wnd{i}_flow"], z[f"wnd{i}_out"]
        assert np.array_equal(co.warping_no_div(x, fl), wo)
        assert np.abs(ops_ref.warping_layer_no_div_ref(torch.from_numpy(x), torch.from_numpy(fl)).numpy() - wo).max() < 1e-6


def test_resize_identities():
    """x1/2 == avg-pool 2^d, x1/4 == mean of samples {4i+1, 4i+2} per axis."""
    x = torch.rand(1, 2, 16, 16, 16)
    half = ops_ref.resize_ref(x, 0.5)
    assert torch.allclose(half, torch.nn.functional.avg_pool3d(x, 2), atol=1e-6)
    quarter = ops_ref.resize_ref(x, 0.25)
    sel = (x[:, :, 1::4] + x[:, :, 2::4]) / 2
    sel = (sel[:, :, :, 1::4] + sel[:, :, :, 2::4]) / 2
    sel = (sel[..., 1::4] + sel[..., 2::4]) / 2
    assert torch.allclose(quarter, sel, atol=1e-6)


def test_metrics_golden():
    """oracle/metrics_ref.py vs the reference's own error.py (fixture written by tests/golden/make_metrics_golden.py, which
    ran calculate_psnr / calculate_ssim of /root/reference/error.py with cv2)."""
    from oracle import metrics_ref as mr
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.npz"))
    for tag in "abcd":
        img1, img2 = z[f"{tag}_img1"], z[f"{tag}_img2"]
        assert abs(mr.calculate_psnr(img1, img2) - float(z[f"{tag}_psnr"])) <= 1e-12 * float(z[f"{tag}_psnr"])
        assert abs(mr.calculate_ssim(img1, img2) - float(z[f"{tag}_ssim"])) <= 1e-10
    assert mr.calculate_psnr(z["a_img1"], z["a_img1"]) == float("inf")
    with pytest.raises(ValueError):
        mr.calculate_ssim(z["a_img1"], z["b_img1"])
    # the training-loop form on [0,1] data is the same quantity on another scale (Flow-3D/train.py:385)
    p01 = mr.psnr_train(z["a_img1"] / 255.0, z["a_img2"] / 255.0)
    assert abs(p01 - float(z["a_psnr"])) <= 1e-5


def test_warp_backward_golden():
    """Backward of warp (autograd under Flow-*/model/warplayer.py): torch-autograd oracle and the numpy restatement
    both replay the vectors tests/golden/make_warp_bwd_golden.py took from autograd through the reference's warp."""
    from oracle.warp_bwd_ref import warp_bwd
    z = np.load(os.path.join(G, "warp_bwd.npz"))
    n = 0
    for name, (src, flow, gout, gsrc, gflow) in _cases(z, ("src", "flow", "gout", "gsrc", "gflow")):
        nd = flow.shape[1]
        a, b = torch.from_numpy(src).requires_grad_(), torch.from_numpy(flow).requires_grad_()
        (ops_ref.warp2d_ref if nd == 2 else ops_ref.warp3d_ref)(a, b).backward(torch.from_numpy(gout))
        assert np.abs(a.grad.numpy() - gsrc).max() <= 1e-5 * max(1.0, np.abs(gsrc).max()), name
        assert np.abs(b.grad.numpy() - gflow).max() <= 1e-5 * max(1.0, np.abs(gflow).max()), name
        gs, gf = warp_bwd(src, flow, gout)
        assert np.abs(gs - gsrc).max() <= 1e-5 * max(1.0, np.abs(gsrc).max()), name
        assert np.abs(gf - gflow).max() <= 1e-5 * max(1.0, np.abs(gflow).max()), name
        if "far" in name or "edge" in name:      # clipped coordinates carry no flow gradient
            assert (gflow == 0).mean() > (0.9 if "edge" in name else 0.5), name
        n += 1
    assert n == 16


def test_upflow_backward_golden():
    """Backward of a10 / a11: autograd through the oracle restatements replays the vectors taken from autograd through the
    reference's pwc_modules (tests/golden/make_upflow_bwd_golden.py)."""
    z = np.load(os.path.join(G, "upflow_bwd.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(4):
            fl, go, gin = (torch.from_numpy(z[f"ups{i}_{s}"]) for s in ("in", "gout", "gin"))
            a = fl.clone().requires_grad_()
            ops_ref.upsample2d_flow_as_ref(a, go.shape[2], go.shape[3]).backward(go)
            assert (a.grad - gin).abs().max() <= 1e-5 * max(1.0, float(gin.abs().max()))
            x, f, go, gx, gf = (torch.from_numpy(z[f"wnd{i}_{s}"]) for s in ("x", "flow", "gout", "gx", "gflow"))
            a, b = x.clone().requires_grad_(), f.clone().requires_grad_()
            ops_ref.warping_layer_no_div_ref(a, b).backward(go)
            assert (a.grad - gx).abs().max() <= 1e-5 * max(1.0, float(gx.abs().max()))
            assert (b.grad - gf).abs().max() <= 1e-5 * max(1.0, float(gf.abs().max()))


def test_adamw_golden():
    """Optimizer step of the training loop: the numpy restatement replays torch.optim.AdamW as driven by the reference
    (tests/golden/make_adamw_golden.py)."""
    from oracle.adamw_ref import adamw_step
    z = np.load(os.path.join(G, "adamw.npz"))
    lrs = z["lrs"]
    for i in range(5):
        p, m, v = z[f"p0_{i}"], np.zeros_like(z[f"p0_{i}"]), np.zeros_like(z[f"p0_{i}"])
        for t, lr in enumerate(lrs, start=1):
            p, m, v = adamw_step(p, z[f"g{t}_{i}"], m, v, t, float(lr))
            assert np.abs(p - z[f"p{t}_{i}"]).max() <= 1e-6 * max(1e-30, np.abs(z[f"p{t}_{i}"]).max()), (i, t)


@pytest.mark.parametrize("nd,sp", [(2, (32, 48)), (3, (16, 16, 32))])
def test_refine_oracle_replays_reference_golden(nd, sp):
    """oracle/refine_ref.py from the generator's seeds reproduces the sums recorded while it was bit-exact against the reference."""
    from oracle.refine_ref import ContextnetRef, UnetRef
    gold = np.load(os.path.join(G, "refine.npz"))
    torch.manual_seed(77)
    oc, ou = ContextnetRef(nd), UnetRef(nd)
    g = torch.Generator().manual_seed(5)
    cin = 1 if nd == 2 else 3
    x = torch.rand((2, cin) + sp, generator=g)
    flow = torch.randn((2, nd) + sp, generator=g) * 2
    with torch.no_grad():
        fo = oc(x, flow)
        parts = torch.rand((2, 9 if nd == 2 else 17) + sp, generator=g)
        args = (parts[:, :cin], parts[:, cin:2 * cin], parts[:, 2 * cin:3 * cin], parts[:, 3 * cin:4 * cin], parts[:, 4 * cin:4 * cin + 1],
                parts[:, 4 * cin + 1:])
        yo = ou(*args, fo, oc(x.flip(0), flow))
    assert np.allclose([float(f.double().sum()) for f in fo], gold[f"nd{nd}_ctx_sum"], rtol=1e-5, atol=1e-3)
    assert abs(float(yo.double().sum()) - float(gold[f"nd{nd}_unet_sum"])) <= 1e-5 * abs(float(gold[f"nd{nd}_unet_sum"]))
    assert np.allclose(yo.flatten()[:16].numpy(), gold[f"nd{nd}_unet_head"], atol=1e-6)
