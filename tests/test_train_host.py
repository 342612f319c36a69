"""CPU tests of the training tier's host logic (no kernels are launched):

* the tap-form layers opticalflowscivis_b200/train.py builds for the INPUT gradient of every layer type equal torch autograd,
* the tap-form weight gradient maps back to the reference's parameter layouts,
* the whole backward wiring of `_BlockFn` / `ifnet_forward_train` (residual pairs, merged heads, parameter order, teacher block,
  distillation) reproduces the oracle's gradients when the four kernel wrappers are replaced by torch evaluators of the SAME tap
  forms in fp32,
* oracle/train_ref.py (`Model.update`) replays the golden losses / gradient norms / parameter deltas pinned against the reference.
"""
import contextlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from opticalflowscivis_b200 import ifnet, train
from oracle.ifnet_ref import IFNetRef
from oracle.ops_ref import warp2d_ref, warp3d_ref
from oracle.train_ref import TrainerRef, training_triplet
from tap_eval import run_layer

HERE = os.path.dirname(os.path.abspath(__file__))


def _cl(x):
    if x.dim() == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def _pad_c(x, c):
    return F.pad(x, (0, c - x.shape[-1]))


def _uncl(y, c, nd):
    y = y[..., :c].permute(0, 4, 1, 2, 3)
    return y[:, :, 0] if nd == 2 else y


# ---- torch evaluators of the library entry points (fp32, CPU) -----------------------------------------------------------
def wgrad_eval(d, x, gy):
    """ofsv_conv_wgrad_bf16's contract (include/ofsv.h) with torch indexing."""
    x, gy = x.float(), gy.float()
    n, di, hi, wi, _ = x.shape
    T = d.nphase * d.ntaps
    dw = torch.zeros(T, d.Cin_s, d.Cout_w)
    zs, ys, xs = torch.arange(d.Do), torch.arange(d.Ho), torch.arange(d.Wo)
    for ph in range(d.nphase):
        pz, py, px = (ph >> 2) & 1, (ph >> 1) & 1, ph & 1
        g = gy[:, zs * d.out_stride + pz][:, :, ys * d.out_stride + py][:, :, :, xs * d.out_stride + px][..., :d.Cout_w]
        for t in range(d.ntaps):
            oz, oy, ox = (int(d.tap_off[ph * d.ntaps + t][k]) for k in range(3))
            iz, iy, ix = zs * d.in_stride + oz, ys * d.in_stride + oy, xs * d.in_stride + ox
            vz, vy, vx = (iz >= 0) & (iz < di), (iy >= 0) & (iy < hi), (ix >= 0) & (ix < wi)
            xv = x[:, iz.clamp(0, di - 1)][:, :, iy.clamp(0, hi - 1)][:, :, :, ix.clamp(0, wi - 1)]
            valid = (vz.view(-1, 1, 1) & vy.view(1, -1, 1) & vx.view(1, 1, -1)).view(1, d.Do, d.Ho, d.Wo, 1)
            dw[ph * d.ntaps + t] = torch.einsum("ndhwc,ndhwo->co", xv * valid, g)
    return dw


def prelu_bias_bwd_eval(gy, y, slope):
    cs = gy.shape[-1]
    g = gy.float().reshape(-1, cs)
    if slope is None:
        return gy, g.sum(0), None
    yy = y.float().reshape(-1, cs)
    pos = yy > 0
    gp = torch.where(pos, g, g * slope)
    ds = torch.where(pos, torch.zeros_like(g), g * (yy / slope)).sum(0)
    return gp.reshape(gy.shape).to(gy.dtype), gp.sum(0), ds


def _run_layer_eval(lay, d, x, res, y):
    out = run_layer(lay, x.float(), None if res is None else res.float())
    y.copy_(out.reshape(y.shape))
    return y


def _to_cl16_eval(x, nd):
    n, c = x.shape[:2]
    out = torch.zeros((n,) + ((1,) if nd == 2 else ()) + tuple(x.shape[2:]) + (16,))
    (out[:, 0] if nd == 2 else out)[..., :c] = x.permute(0, 2, 3, 1) if nd == 2 else x.permute(0, 2, 3, 4, 1)
    return out


def _from_cl_eval(y, c, nd):
    if nd == 2:
        return y[:, 0, :, :, :c].permute(0, 3, 1, 2).float().contiguous()
    return y[..., :c].permute(0, 4, 1, 2, 3).float().contiguous()


@pytest.fixture
def cpu_engine(monkeypatch):
    monkeypatch.setattr(train, "_require_cuda", lambda t, name: t)
    monkeypatch.setattr(train, "_to_cl16", _to_cl16_eval)
    monkeypatch.setattr(train, "_from_cl", _from_cl_eval)
    monkeypatch.setattr(train, "_ACT_DTYPE", torch.float32)
    monkeypatch.setattr(train, "_run_layer", _run_layer_eval)
    monkeypatch.setattr(train, "conv_wgrad", wgrad_eval)
    monkeypatch.setattr(train, "prelu_bias_bwd", prelu_bias_bwd_eval)
    monkeypatch.setattr(train.ops, "_on", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(train, "_warp_fn", lambda nd: warp2d_ref if nd == 2 else warp3d_ref)


# ---- input-gradient layers ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nd", [2, 3])
def test_dgrad_tap_forms_match_autograd(nd):
    torch.manual_seed(3)
    conv, convT = (F.conv2d, F.conv_transpose2d) if nd == 2 else (F.conv3d, F.conv_transpose3d)
    s = 8
    # Conv(3, 1, 1)
    w = torch.randn((16, 16) + (3,) * nd) * 0.2
    x = torch.randn((2, 16) + (s,) * nd, requires_grad=True)
    gy = torch.randn((2, 16) + (s,) * nd)
    (gx,) = torch.autograd.grad(conv(x, w, padding=1), x, gy)
    lay = train._conv_layer(nd, w, 3, 1, 1, 16, mirror=True)
    assert torch.allclose(_uncl(run_layer(lay, _cl(gy)), 16, nd), gx, atol=1e-4)
    # Conv(k0, 2, 1): 2-D k = 3, 3-D k = 4
    k0 = 3 if nd == 2 else 4
    w = torch.randn((32, 11) + (k0,) * nd) * 0.2
    x = torch.randn((2, 11) + (s,) * nd, requires_grad=True)
    y = conv(x, w, stride=2, padding=1)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    lay = train._convT_phase_layer(nd, w, k0, 16)
    assert lay.nphase == 2 ** nd and lay.out_stride == 2
    assert torch.allclose(_uncl(run_layer(lay, _cl(gy)), 11, nd), gx, atol=1e-4)
    # ConvTranspose(4, 2, 1)
    w = torch.randn((32, 7) + (4,) * nd) * 0.2
    x = torch.randn((2, 32) + (s // 2,) * nd, requires_grad=True)
    y = convT(x, w, stride=2, padding=1)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    wpad = torch.zeros((32, 16) + (4,) * nd)
    wpad[:, :7] = w
    lay = train._conv_layer(nd, wpad, 4, 2, 1, 32)
    assert torch.allclose(_uncl(run_layer(lay, _pad_c(_cl(gy), 16)), 32, nd), gx, atol=1e-4)


@pytest.mark.parametrize("nd", [2, 3])
def test_wgrad_tap_form_maps_back_to_parameters(nd):
    torch.manual_seed(4)
    blk = ifnet.IFBlock(nd, 5 + 2 * nd, c=32)
    tb = train._TrainBlock(blk)
    tb.refresh()
    conv, convT = (F.conv2d, F.conv_transpose2d) if nd == 2 else (F.conv3d, F.conv_transpose3d)
    s = 8
    # conv0.0 (strided) through ifnet._pack_conv's tap order
    m = blk.conv0[0][0]
    x = torch.randn((2, m.cin) + (s,) * nd)
    w = m.weight.detach().clone().requires_grad_(True)
    y = conv(x, w, stride=2, padding=1)
    gy = torch.randn_like(y)
    (gw,) = torch.autograd.grad(y, w, gy)
    lay = tb.fwd[0]
    d, _ = lay.desc(2, ((1,) if nd == 2 else ()) + (s,) * nd, 0, has_residual=False)
    got = tb.conv_weight_grad(wgrad_eval(d, _pad_c(_cl(x), lay.cin_s), _pad_c(_cl(gy), lay.cout_w)), m)
    assert torch.allclose(got, gw, atol=1e-3, rtol=1e-4)
    # merged ConvT conv1.0 ‖ conv2.0 through ifnet._pack_convT's (parity, choice) order
    c = blk.c
    wm = torch.cat([blk.conv1[0].weight, blk.conv2[0].weight], 1).detach().clone().requires_grad_(True)
    x = torch.randn((2, c) + (s // 2,) * nd)
    y = convT(x, wm, stride=2, padding=1)
    gy = torch.randn_like(y)
    (gw,) = torch.autograd.grad(y, wm, gy)
    lay = tb.fwd[10]
    d, _ = lay.desc(2, ((1,) if nd == 2 else ()) + (s // 2,) * nd, 0, has_residual=False)
    got = tb.convT_weight_grad(wgrad_eval(d, _cl(x), _cl(gy)), c, c)
    assert torch.allclose(got, gw, atol=1e-3, rtol=1e-4)


# ---- the whole backward wiring ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nd", [2, 3])
def test_block_function_gradients_match_oracle_block(nd, cpu_engine):
    torch.manual_seed(11)
    cin = 5 + 2 * nd
    ref = IFNetRef(nd).block2
    blk = ifnet.IFBlock(nd, cin, c=64)
    blk.load_state_dict(ref.state_dict())
    s = 16
    x = torch.randn((2, cin) + (s,) * nd)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    tb = train._TrainBlock(blk)
    head = train._BlockFn.apply(xa, tb, False, *blk.parameters())
    hr = ref.conv0(xb)
    for i in range(4):
        hr = getattr(ref, f"convblock{i}")(hr) + hr
    head_ref = torch.cat((ref.conv1(hr), ref.conv2(hr)), 1)
    assert torch.allclose(head, head_ref, atol=2e-4)
    g = torch.randn_like(head_ref)
    head.backward(g)
    head_ref.backward(g)
    assert torch.allclose(xa.grad, xb.grad, atol=1e-4, rtol=1e-3)
    for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, k
        scale = max(1e-3, q.grad.abs().max().item())
        assert (p.grad - q.grad).abs().max().item() <= 2e-4 * scale, (k, (p.grad - q.grad).abs().max().item(), scale)


@pytest.mark.parametrize("nd", [2, 3])
def test_training_forward_and_gradients_match_oracle(nd, cpu_engine):
    """ifnet_forward_train + the update losses on the CPU evaluators == oracle/train_ref.py (itself pinned against the reference)."""
    torch.manual_seed(1234)
    ref = IFNetRef(nd)
    net = ifnet.IFNet(nd)
    net.load_state_dict(ref.state_dict())
    size = 32
    img0, img1, gt = training_triplet(nd, 1, size)
    x = torch.cat((img0, img1, gt), 1)
    flow, mask, merged, flow_t, merged_t, ld = train.ifnet_forward_train(net, x)
    fr, mr, mgr, ftr, mtr, ldr = ref.forward_train(x)
    assert torch.allclose(flow[2], fr[2], atol=1e-4)
    assert torch.allclose(merged[2], mgr[2], atol=1e-4)
    assert torch.allclose(flow_t, ftr, atol=1e-4) and torch.allclose(merged_t, mtr, atol=1e-4)
    assert abs(float(ld.detach()) - float(ldr.detach())) <= 1e-4 * max(1.0, abs(float(ldr.detach())))
    loss = F.l1_loss(merged[2], gt) + F.l1_loss(merged_t, gt) + 0.1 * ld
    loss_r = F.l1_loss(mgr[2], gt) + F.l1_loss(mtr, gt) + 0.1 * ldr
    loss.backward()
    loss_r.backward()
    bad = []
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, k
        scale = max(1e-7, q.grad.abs().max().item())
        err = (p.grad - q.grad).abs().max().item()
        if err > 5e-3 * scale:
            bad.append((k, err, scale))
    assert not bad, bad[:5]


# ---- the oracle's update against the golden record of the reference's ------------------------------------------------------------
@pytest.mark.parametrize("nd,n,size", [(3, 2, 32), (2, 2, 64)])
def test_oracle_update_replays_reference_golden(nd, n, size):
    gold = np.load(os.path.join(HERE, "golden", "update.npz"))
    torch.manual_seed(1234)
    tr = TrainerRef(nd)
    w1 = sum(v.double().abs().sum().item() for v in tr.flownet.state_dict().values())
    assert abs(w1 - float(gold[f"nd{nd}_w1"])) <= 1e-9 * w1          # same seeded initialisation as the reference's Model()
    p0 = {k: v.clone() for k, v in tr.flownet.state_dict().items()}
    img0, img1, gt = training_triplet(nd, n, size)
    imgs = torch.cat((img0, img1), 1)
    for step in range(3):
        _, info = tr.update(imgs, gt, learning_rate=1e-4, training=True)
        for key in ("loss_l1", "loss_tea", "loss_distill", "loss_G"):
            a, b = float(info[key]), float(gold[f"nd{nd}_step{step}_{key}"])
            assert abs(a - b) <= 2e-6 * max(1.0, abs(b)), (step, key, a, b)
        if step == 0:
            gn = np.array([p.grad.double().norm().item() for _, p in tr.flownet.named_parameters()])
            assert np.allclose(gn, gold[f"nd{nd}_gradnorm"], rtol=1e-4, atol=1e-10)
    names = [k for k, _ in tr.flownet.named_parameters()]
    sd = tr.flownet.state_dict()
    dn = np.array([(sd[k] - p0[k]).double().norm().item() for k in names])
    assert np.allclose(dn, gold[f"nd{nd}_deltanorm"], rtol=1e-3, atol=1e-9)


@pytest.mark.parametrize("nd,c", [(3, 64), (2, 96)])
def test_train_block_refresh_equals_rebuild(nd, c):
    """The per-step weight refresh (index_select + strided copy per parameter) gives exactly the tap-form tensors of a fresh build."""
    torch.manual_seed(5)
    blk = ifnet.IFBlock(nd, 6 + 2 * nd, c=c)
    tb = train._TrainBlock(blk)
    tb.refresh()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    tb.refresh()
    fresh = train._TrainBlock(blk)
    fresh.refresh()
    for kind in ("fwd", "dgrad"):
        for li, (a, b) in enumerate(zip(getattr(tb, kind), getattr(fresh, kind))):
            assert torch.equal(a.w_simt, b.w_simt), (kind, li)
            assert torch.equal(a.bias, b.bias), (kind, li)
            assert (a.prelu is None) == (b.prelu is None) and (a.prelu is None or torch.equal(a.prelu, b.prelu)), (kind, li)
            assert a.taps == b.taps and not a._packed
