"""GPU parity tests of the training tier: the backward kernels through the C ABI against fp32 torch evaluators of
the same contracts, one IFBlock's gradients against the oracle block, and `Model.update` against oracle/train_ref.py (pinned
against the reference's own `update` by tests/golden/make_update_golden.py).

Tolerances: the kernels take bf16 operands and accumulate in fp32, so against an fp32 evaluation of the SAME bf16-rounded operands
they agree to summation-order noise (1e-3 relative to the tensor's largest entry); against the fp32 oracle the bf16 storage of
activations and gradients shows up: per-tensor gradient cosine >= 0.99 (>= 0.98 for the whole net in one vector is never needed —
the measured values are printed), losses within 2e-3 relative."""
import os

import numpy as np
import pytest
import torch

from test_train_host import prelu_bias_bwd_eval, wgrad_eval

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


@pytest.mark.parametrize("cs,rows", [(16, 1000), (32, 4096), (48, 777), (64, 20000), (96, 333), (128, 5000)])
def test_prelu_bias_bwd_kernel(cs, rows):
    from opticalflowscivis_b200 import train
    torch.manual_seed(cs)
    dev = _dev()
    gy = torch.randn(rows, cs, device=dev).bfloat16()
    y = torch.randn(rows, cs, device=dev).bfloat16()
    slope = torch.rand(cs, device=dev) * 0.4 + 0.05
    gp, db, ds = train.prelu_bias_bwd(gy, y, slope)
    gp_r, db_r, ds_r = prelu_bias_bwd_eval(gy.cpu(), y.cpu(), slope.cpu())
    assert torch.equal(gp.cpu(), gp_r)                                   # element-wise: identical rounding
    assert torch.allclose(db.cpu(), db_r, atol=1e-3 * rows ** 0.5, rtol=1e-4)
    assert torch.allclose(ds.cpu(), ds_r, atol=1e-3 * rows ** 0.5 * 10, rtol=1e-4)
    gp2, db2, ds2 = train.prelu_bias_bwd(gy, None, None)                 # layer without activation
    assert gp2 is gy and ds2 is None
    assert torch.allclose(db2.cpu(), gy.float().sum(0).cpu(), atol=1e-3 * rows ** 0.5, rtol=1e-4)
    a, b, c = train.prelu_bias_bwd(gy, y, slope)                         # deterministic
    assert torch.equal(a, gp) and torch.equal(b, db) and torch.equal(c, ds)


def _layers(nd):
    """(name, forward _Layer, logical input dims) for every layer family and tile shape of the training path."""
    from opticalflowscivis_b200 import ifnet
    out = []
    for c in ((64, 128) if nd == 3 else (64, 96)):
        blk = ifnet.IFBlock(nd, 5 + 2 * nd, c=c)
        L = blk.layers()
        s = 24 if c == 64 else 16          # 24: grids of 12 / 6 positions per axis = partial bricks of the brick-window kernel
        out += [(f"c{c}.conv0.0", L[0], s), (f"c{c}.conv0.1", L[1], s // 2), (f"c{c}.convblock", L[2], s // 4),
                (f"c{c}.convT", L[10], s // 4), (f"c{c}.heads", L[11], s // 2)]
    return out


@pytest.mark.parametrize("brick", [1, 0])
@pytest.mark.parametrize("nd", [2, 3])
def test_conv_wgrad_kernel_vs_fp32_evaluator(nd, brick):
    """brick = 1: the brick-window kernel wherever its windows fit; 0: the per-tap / tap-group kernels (ofsv_set_tuning('wgrad_brick', v);
    the default policy -1 mixes them per layer and is what every other test runs)."""
    from opticalflowscivis_b200 import _C, ops, train
    torch.manual_seed(nd)
    dev = _dev()
    n = 2
    ops.set_tuning("wgrad_brick", brick)
    try:
        _wgrad_cases(nd, n, dev, _C, train)
    finally:
        ops.set_tuning("wgrad_brick", -1)


def _wgrad_cases(nd, n, dev, _C, train):
    for name, lay, s in _layers(nd):
        in_sp = ((1,) if nd == 2 else ()) + (s,) * nd
        d, osp = lay.desc(n, in_sp, _C.BF16, has_residual=False)
        x = torch.randn((n,) + (((1,) + (s,) * 2) if nd == 2 else (s,) * 3) + (lay.cin_s,), device=dev).bfloat16()
        gs = 16 if lay.out_f32 else lay.cout_s
        gy = torch.randn((n,) + tuple(osp) + (gs,), device=dev).bfloat16()
        dw = train.conv_wgrad(d, x, gy)
        ref = wgrad_eval(d, x.cpu(), gy.cpu())
        scale = ref.abs().max().item()
        err = (dw.cpu() - ref).abs().max().item()
        assert err <= 2e-3 * scale, (name, err, scale)
        assert torch.equal(train.conv_wgrad(d, x, gy), dw), name          # fixed-order K split: deterministic


@pytest.mark.parametrize("nd,c", [(3, 64), (3, 128), (2, 96)])
def test_block_function_gradients_vs_oracle_block(nd, c):
    from opticalflowscivis_b200 import ifnet, train
    from oracle.ifnet_ref import IFBlockRef
    torch.manual_seed(7)
    dev = _dev()
    cin = 5 + 2 * nd
    ref = IFBlockRef(nd, cin, c).to(dev)
    blk = ifnet.IFBlock(nd, cin, c=c).to(dev)
    blk.load_state_dict(ref.state_dict())
    s = 32 if nd == 3 else 64
    x = torch.randn((2, cin) + (s,) * nd, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    head = train._BlockFn.apply(xa, train._TrainBlock(blk), False, *blk.parameters())
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        hr = ref.conv0(xb)
        for i in range(4):
            hr = getattr(ref, f"convblock{i}")(hr) + hr
        head_ref = torch.cat((ref.conv1(hr), ref.conv2(hr)), 1)
        g = torch.randn_like(head_ref)
        head_ref.backward(g)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    head.backward(g)
    assert _cos(head, head_ref) >= 0.9995
    assert _cos(xa.grad, xb.grad) >= 0.995, _cos(xa.grad, xb.grad)
    worst = min((_cos(p.grad, q.grad), k) for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()))
    print("block gradient cosine, worst tensor:", worst)
    assert worst[0] >= 0.99, worst
    for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        r = float(p.grad.double().norm() / (q.grad.double().norm() + 1e-300))
        assert 0.97 <= r <= 1.03, (k, r)


@pytest.mark.parametrize("nd,n,size", [(3, 2, 64), (2, 4, 128)])
def test_model_update_vs_oracle(nd, n, size):
    """Three `Model.update` steps against the oracle's (fp32 eager on the same GPU, TF32 off): losses, first-step gradients,
    parameter movement.  AdamW's first steps move every weight by ~lr * sign(gradient), so the parameter-delta check is a
    gradient-SIGN check weighted by nothing: it is stated as a cosine over all parameters."""
    from opticalflowscivis_b200.rife import Model2D, Model3D
    from oracle.train_ref import TrainerRef, training_triplet
    dev = _dev()
    torch.manual_seed(1234)
    orc = TrainerRef(nd)
    orc.flownet.to(dev)
    orc.optimG = torch.optim.AdamW(orc.flownet.parameters(), lr=1e-6, weight_decay=1e-3)
    model = (Model3D if nd == 3 else Model2D)(local_rank=-1)
    model.flownet.load_state_dict(orc.flownet.state_dict())
    p0 = [p.detach().clone() for p in model.flownet.parameters()]
    img0, img1, gt = (t.to(dev) for t in training_triplet(nd, n, size))
    imgs = torch.cat((img0, img1), 1)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for step in range(3):
            mg_r, ir = orc.update(imgs, gt, learning_rate=1e-4, training=True)
            if step == 0:
                g_ref = [p.grad.detach().clone() for p in orc.flownet.parameters()]
            mg, info = model.update(imgs, gt, learning_rate=1e-4, training=True) if nd == 3 else \
                model.update(imgs, gt, "droplet2d", learning_rate=1e-4, training=True)
            if step == 0:
                g_mine = [p.grad.detach().clone() for p in model.flownet.parameters()]
            for key in ("loss_l1", "loss_tea", "loss_G"):
                a, b = float(info[key]), float(ir[key])
                assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (step, key, a, b)
            a, b = float(info["loss_distill"]), float(ir["loss_distill"])
            assert abs(a - b) <= 2e-2 * max(abs(b), 1e-3), (step, "loss_distill", a, b)
            psnr = -10 * torch.log10(((mg - mg_r.detach()) ** 2).mean()).item()
            assert psnr >= 50.0, (step, psnr)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    names = [k for k, _ in model.flownet.named_parameters()]
    cosines = sorted((_cos(a, b), k) for a, b, k in zip(g_mine, g_ref, names))
    allcos = _cos(torch.cat([g.flatten() for g in g_mine]), torch.cat([g.flatten() for g in g_ref]))
    print(f"update nd={nd}: first-step gradient cosine over all parameters {allcos:.5f}; worst tensors {cosines[:4]}")
    assert allcos >= 0.995
    weights = [(c, k) for c, k in cosines if k.endswith("0.weight") or k.endswith("2.weight")]
    assert min(weights)[0] >= 0.98, min(weights)
    d_mine = torch.cat([(p.detach() - q).flatten() for p, q in zip(model.flownet.parameters(), p0)])
    d_ref = torch.cat([(p.detach() - q).flatten() for p, q in zip(orc.flownet.parameters(), p0)])
    dcos = _cos(d_mine, d_ref)
    print(f"update nd={nd}: parameter-delta cosine after 3 steps {dcos:.4f}; |delta| mine {d_mine.norm():.4e} ref {d_ref.norm():.4e}")
    assert dcos >= 0.9
    assert abs(float(d_mine.norm() / d_ref.norm()) - 1) <= 0.05
    # evaluation mode: no parameter change, teacher outputs aliased to the student's (RIFE.py:260-262)
    before = [p.detach().clone() for p in model.flownet.parameters()]
    mg, info = model.update(imgs, gt, learning_rate=1e-4, training=False) if nd == 3 else \
        model.update(imgs, gt, "droplet2d", learning_rate=1e-4, training=False)
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, model.flownet.parameters()))
    assert torch.equal(info["merged_tea"], mg)
    # and inference sees the updated weights (packed tap forms are keyed on parameter versions)
    out = model.inference(img0, img1)
    assert torch.isfinite(out[0] if nd == 3 else out[0][2]).all()


@pytest.mark.parametrize("nd,n,size", [(3, 1, 64), (2, 2, 64)])
def test_model_update_graph_mode_equals_eager(nd, n, size):
    """`enable_training_graph()` replays the same kernels: losses and parameters follow the eager run (not bit-for-bit: the warp
    backward accumulates with red.global.add in whatever order the SMs get there)."""
    from opticalflowscivis_b200.rife import Model2D, Model3D
    from oracle.train_ref import training_triplet
    dev = _dev()
    torch.manual_seed(1234)
    M = Model3D if nd == 3 else Model2D
    a, b = M(local_rank=-1), M(local_rank=-1)
    b.flownet.load_state_dict(a.flownet.state_dict())
    b.enable_training_graph()
    img0, img1, gt = (t.to(dev) for t in training_triplet(nd, n, size))
    imgs = torch.cat((img0, img1), 1)
    for step in range(4):
        lr = 1e-4 * (step + 1)                      # the learning rate changes every step (RIFE.py:86-87) — outside the graph
        args = (imgs, gt) if nd == 3 else (imgs, gt, "droplet2d")
        _, ia = a.update(*args, learning_rate=lr, training=True)
        mg, ib = b.update(*args, learning_rate=lr, training=True)
        for key in ("loss_l1", "loss_tea", "loss_distill", "loss_G"):
            x, y = float(ia[key]), float(ib[key])
            assert abs(x - y) <= (5e-3 if key == "loss_distill" else 5e-4) * max(abs(x), 1e-3), (step, key, x, y)
    da = torch.cat([p.detach().flatten() for p in a.flownet.parameters()])
    db = torch.cat([p.detach().flatten() for p in b.flownet.parameters()])
    assert _cos(da, db) >= 0.999999
    assert (da - db).abs().max().item() <= 2.5e-3      # a sign flip of a ~zero gradient moves a weight by 2 * lr per step
    assert len(b._trainer.graphs) == 1


# ------------------------------------------------------------------------------------------------ refinement nets (§8 f.3)
@pytest.mark.parametrize("nd,sp", [(2, (64, 96)), (2, (160, 224)), (3, (32, 32, 48))])
def test_refine_nets_vs_oracle(nd, sp):
    """Contextnet / Unet on the bf16 engines against the fp32 oracle (pinned bit-exact against Flow-{2D,3D}/model/refine.py)."""
    from opticalflowscivis_b200 import refine
    from oracle.refine_ref import ContextnetRef, UnetRef
    dev = _dev()
    torch.manual_seed(77)
    oc, ou = ContextnetRef(nd).to(dev), UnetRef(nd).to(dev)
    mc, mu = refine.Contextnet(nd).to(dev), refine.Unet(nd).to(dev)
    mc.load_state_dict(oc.state_dict())
    mu.load_state_dict(ou.state_dict())
    g = torch.Generator().manual_seed(5)
    cin = 1 if nd == 2 else 3
    x = torch.rand((2, cin) + sp, generator=g).to(dev)
    flow = (torch.randn((2, nd) + sp, generator=g) * 2).to(dev)
    parts = torch.rand((2, 9 if nd == 2 else 17) + sp, generator=g).to(dev)
    args = (parts[:, :cin], parts[:, cin:2 * cin], parts[:, 2 * cin:3 * cin], parts[:, 3 * cin:4 * cin], parts[:, 4 * cin:4 * cin + 1],
            parts[:, 4 * cin + 1:])
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            fo, fo1 = oc(x, flow), oc(x.flip(0), flow)
            yo = ou(*args, fo, fo1)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    fm, fm1 = mc(x, flow), mc(x.flip(0), flow)
    for i, (a, b) in enumerate(zip(fm, fo)):
        assert a.shape == b.shape
        rel = float((a - b).norm() / (b.norm() + 1e-12))
        assert rel <= 1.5e-2, (i, rel)                       # bf16 activations through 2 (i + 1) conv layers
    ym = mu(*args, fm, fm1)
    assert ym.shape == yo.shape
    err = float((ym - yo).abs().max())
    print(f"refine nd={nd} {sp}: Unet output max-abs {err:.2e}")
    assert err <= 1e-2, err                                   # sigmoid output in [0,1]


def test_ifnet2d_refine_switch_vs_oracle():
    """`refine = True` (Flow-2D/model/IFNet.py:32,255-273): merged[2] = clamp(merged[2] + unet(...) * 2 - 1, 0, 1)."""
    import importlib
    mod = importlib.import_module("opticalflowscivis_b200.flow2d.model.IFNet")
    from oracle.ifnet_ref import IFNetRef
    from oracle.ops_ref import warp2d_ref
    from oracle.refine_ref import ContextnetRef, UnetRef, refine_merged_ref
    dev = _dev()
    torch.manual_seed(1234)
    onet, oc, ou = IFNetRef(2).to(dev), ContextnetRef(2).to(dev), UnetRef(2).to(dev)
    mod.refine = True
    try:
        net = mod.IFNet().to(dev)
    finally:
        mod.refine = False
    sd = dict(onet.state_dict())
    sd.update({"contextnet." + k: v for k, v in oc.state_dict().items()})
    sd.update({"unet." + k: v for k, v in ou.state_dict().items()})
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.rand((2, 2, 64, 96), generator=g).to(dev)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            fo, mo, mgo = onet(x, (4, 2, 1))
            w0, w1 = warp2d_ref(x[:, :1], fo[2][:, :2]), warp2d_ref(x[:, 1:2], fo[2][:, 2:4])
            ref = refine_merged_ref(oc, ou, x[:, :1], x[:, 1:2], w0, w1, torch.logit(mo[2]), fo[2], mgo[2])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    flow, mask, merged, *_ = net(x, (4, 2, 1))
    err = float((merged[2] - ref).abs().max())
    psnr = -10 * float(torch.log10(((merged[2] - ref) ** 2).mean()))
    print(f"IFNet2D refine=True: merged[2] max-abs {err:.2e}, PSNR {psnr:.1f} dB vs the fp32 oracle")
    assert err <= 3e-2 and psnr >= 45.0
    assert float((merged[2] - mgo[2]).abs().max()) > 0.01      # the refinement did change the frame (0.034 with this init)


# ------------------------------------------------------------------------------------------------ UPFlow decode (§8 f.2)
@pytest.mark.parametrize("shape", [(2, 196, 4, 13), (4, 128, 8, 26), (2, 64, 32, 104), (2, 32, 64, 208), (1, 7, 9, 11)])
def test_feature_norm_pair_and_torch_warp_vs_oracle(shape):
    """ofsv_feature_norm_pair_f32 (warp + per-plane normalisation, upflow.py:621-640 / :95-138) and ofsv_torch_warp_f32
    (tools.py:1317-1361) against the restatements pinned on the reference (oracle/upflow_ref.py, oracle/ops_ref.py)."""
    from opticalflowscivis_b200 import ops
    from oracle import ops_ref, upflow_ref as ur
    dev = _dev()
    g = torch.Generator().manual_seed(shape[1])
    b, c, h, w = shape
    a = torch.randn(shape, generator=g) * 2 + 0.5
    o = torch.randn(shape, generator=g) * 3 - 1
    flow = torch.randn((b, 2, h, w), generator=g) * 2.5
    w_ref = ops_ref.warping_layer_no_div_ref(o, flow)
    na, nw = ur.normalize_features_ref(a, w_ref)
    ga, gw = ops.feature_norm_pair(a.to(dev), o.to(dev), flow.to(dev))
    assert float((ga.cpu() - na).abs().max()) <= 2e-5
    assert float((gw.cpu() - nw).abs().max()) <= 2e-5 * max(1.0, float(nw.abs().max()))
    na0, no0 = ur.normalize_features_ref(a, o)                              # level 0: no warp
    ga0, go0 = ops.feature_norm_pair(a.to(dev), o.to(dev), None)
    assert float((ga0.cpu() - na0).abs().max()) <= 2e-5 and float((go0.cpu() - no0).abs().max()) <= 2e-5
    tw = ops.torch_warp(o.to(dev), flow.to(dev)).cpu()
    assert float((tw - ur.torch_warp_ref(o, flow)).abs().max()) <= 1e-5
    x, y = ops.feature_norm_pair(a.to(dev), o.to(dev), flow.to(dev))        # deterministic
    assert torch.equal(x, ga) and torch.equal(y, gw)


@pytest.mark.parametrize("tag,sgu", [("train", False), ("sgu", True)])
def test_upflow_net_vs_reference_record(tag, sgu):
    """UPFlowNet.forward_2_frame_v3 on the bf16 engines against the record of the reference's UPFlow_net (tests/golden/upflow_net.npz).
    Per level, teacher-forced (see tests/test_upflow_net_host.py): the error is the bf16 arithmetic of that level's 20 convolutions;
    free-running, the coarse levels are compared directly and the output by its mean error (the reference's `>= 1` warp mask makes
    the fine levels chaotic under ANY change of arithmetic)."""
    from opticalflowscivis_b200.upflow import net as unet
    from oracle import upflow_ref as ur
    from test_upflow_net_host import teacher_forced_level_errors
    dev = _dev()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "upflow_net.npz"))
    net = unet.UPFlowNet(if_sgu_upsample=sgu)
    net.load_state_dict(ur.deterministic_state({k: tuple(v.shape) for k, v in net.state_dict().items()}))
    net.to(dev)
    im1, im2 = (t.to(dev) for t in ur.smooth_pair(1, 128, 192))
    errs = teacher_forced_level_errors(net, gold, tag, im1, im2)
    mags = [float(np.abs(gold[f"{tag}_lvl{4 - l}_f"]).max()) for l in range(5)] + [float(np.abs(gold[f"{tag}_flow_f_sub4"]).max())]
    print(f"UPFlowNet ({tag}) teacher-forced max-abs error per level {['%.1e' % e for e in errs]} px; max |flow| {['%.2f' % m for m in mags]} px")
    for e, m in zip(errs, mags):
        # north_star: flow within 1e-2 px with bf16 convs (+ 3 % of |flow|).  With the self-guided up-sampling the warp INSIDE a level
        # uses a flow that went through the bf16 sgu net, so a few `>= 1` mask pixels flip there as well: twice the allowance
        assert e <= (2 if sgu else 1) * (1e-2 + 0.03 * m), (errs, mags)
    ff, fb, flows = net.forward_2_frame_v3(im1, im2)
    assert ff.shape == (1, 2, 128, 192) and len(flows) == 5
    for i in (4, 3, 2):
        assert float(np.abs(flows[i][0].cpu().numpy() - gold[f"{tag}_lvl{i}_f"]).max()) <= 1e-2
    ref = gold[f"{tag}_flow_f_sub4"]
    err = float(np.abs(ff[:, :, ::4, ::4].cpu().numpy() - ref).mean())
    print(f"UPFlowNet ({tag}) free-running: mean |flow error| {err:.3e} px at mean |flow| {float(np.abs(ref).mean()):.3f} px")
    assert err <= 0.1 * float(np.abs(ref).mean())
    if not sgu:                                                  # CUDA-graph replay: same kernels, same results
        net.enable_cuda_graphs()
        for _ in range(2):
            gf, gb, gflows = net.forward_2_frame_v3(im1, im2)
        assert torch.equal(gf, ff) and torch.equal(gb, fb) and torch.equal(gflows[0][0], flows[0][0])
        net.enable_cuda_graphs(False)
    of, ob = unet.occ_check(ff, fb)
    assert of.shape == (1, 1, 128, 192) and float(of.min()) >= 0 and float(of.max()) <= 1
    rf, rb = ur.occ_check_ref(ff.cpu(), fb.cpu())
    assert float((of.cpu() - rf).abs().mean()) <= 2e-3           # thresholded: a few pixels may sit on the threshold


@pytest.mark.parametrize("nd,c", [(3, 64), (3, 128), (2, 96)])
def test_train_block_refresh_kernel_equals_rebuild(nd, c):
    """ofsv_conv_refresh_tapform (one launch for the 24 layers of a block) == the tap forms of a fresh build, exactly."""
    from opticalflowscivis_b200 import ifnet, train
    torch.manual_seed(5)
    blk = ifnet.IFBlock(nd, 6 + 2 * nd, c=c).to(_dev())
    tb = train._TrainBlock(blk)
    tb.refresh()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    tb.refresh()
    assert tb._table is not None
    fresh = train._TrainBlock(blk)
    fresh.refresh()
    for kind in ("fwd", "dgrad"):
        for li, (a, b) in enumerate(zip(getattr(tb, kind), getattr(fresh, kind))):
            assert torch.equal(a.w_simt, b.w_simt), (kind, li)
            assert torch.equal(a.bias, b.bias), (kind, li)
            assert (a.prelu is None) == (b.prelu is None) and (a.prelu is None or torch.equal(a.prelu, b.prelu)), (kind, li)


@pytest.mark.parametrize("nd,sp", [(3, (32, 48, 64)), (2, (96, 160))])
def test_training_gradients_noncubic_vs_oracle(nd, sp):
    """Forward with the teacher block + backward on a non-cubic volume / non-square frame (the 3-D warp then runs on the gather
    kernel, the weight-gradient bricks and tiles are partial along some axes): first-step gradients against the fp32 oracle's."""
    from opticalflowscivis_b200 import ifnet, train
    from oracle.ifnet_ref import IFNetRef
    dev = _dev()
    torch.manual_seed(1234)
    ref = IFNetRef(nd).to(dev)
    net = ifnet.IFNet(nd).to(dev)
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(3)
    base = torch.rand((2, 1) + tuple(s + 8 for s in sp), generator=g)
    pool = torch.nn.functional.avg_pool3d if nd == 3 else torch.nn.functional.avg_pool2d
    base = pool(base, 5, 1, 2).to(dev)
    crop = lambda o: base[(slice(None), slice(None)) + tuple(slice(4, 4 + s) for s in sp[:-1]) + (slice(o, o + sp[-1]),)].contiguous()  # noqa: E731
    x = torch.cat((crop(2), crop(6), crop(4)), 1)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fr, mr, mgr, ftr, mtr, ldr = ref.forward_train(x)
        loss_r = (mgr[2] - x[:, 2:3]).abs().mean() + (mtr - x[:, 2:3]).abs().mean() + 0.1 * ldr
        loss_r.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    flow, mask, merged, flow_t, merged_t, ld = train.ifnet_forward_train(net, x)
    loss = (merged[2] - x[:, 2:3]).abs().mean() + (merged_t - x[:, 2:3]).abs().mean() + 0.1 * ld
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_r.detach())) <= 2e-3 * abs(float(loss_r.detach()))
    ga = torch.cat([p.grad.flatten() for p in net.parameters()])
    gb = torch.cat([p.grad.flatten() for p in ref.parameters()])
    c = _cos(ga, gb)
    worst = min((_cos(p.grad, q.grad), k) for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()) if k.endswith("0.weight") or k.endswith("2.weight"))
    print(f"non-cubic {sp}: gradient cosine {c:.5f}, worst weight tensor {worst}")
    assert c >= 0.999 and worst[0] >= 0.98


@pytest.mark.parametrize("shape,chans,cs", [((2, 37, 45), (81, 32, 2), 128), ((1, 5, 6, 7), (11,), 16), ((3, 64, 208), (196,), 208),
                                            ((2, 8, 8), (7, 9, 3, 5), 32), ((1, 16, 16, 16), (64,), 64)])
def test_pack_unpack_nhwc_kernels(shape, chans, cs):
    """ofsv_pack_nhwc_bf16 (channel concatenation + zero padding + cast, one launch) and ofsv_unpack_nhwc_f32 against torch."""
    from opticalflowscivis_b200 import ops
    dev = _dev()
    torch.manual_seed(sum(chans))
    n, sp = shape[0], shape[1:]
    nd = len(sp)
    srcs = [torch.randn((n, c) + sp, device=dev) for c in chans]
    got = ops.pack_nhwc(srcs, cs)
    cat = torch.cat(srcs, 1)
    ref = torch.zeros((n,) + ((1,) + sp if nd == 2 else sp) + (cs,), device=dev, dtype=torch.bfloat16)
    (ref[:, 0] if nd == 2 else ref)[..., :cat.shape[1]] = cat.permute(0, 2, 3, 1) if nd == 2 else cat.permute(0, 2, 3, 4, 1)
    assert got.shape == ref.shape and torch.equal(got, ref)
    c = cat.shape[1]
    back = ops.unpack_nhwc(got, c, nd)
    assert back.shape == cat.shape and torch.equal(back, cat.bfloat16().float())
    with pytest.raises(TypeError):
        ops.pack_nhwc([srcs[0].cpu()], cs)


@pytest.mark.parametrize("nd,sp,scale", [(3, (16, 32, 48), 4), (3, (16, 16, 32), 2), (3, (8, 16, 16), 1), (2, (64, 96), 4), (2, (32, 48), 2),
                                         (2, (16, 32), 1)])
def test_pack_input_and_head_upsample_backward_vs_autograd(nd, sp, scale):
    """ofsv_pack_block_input_bwd / ofsv_head_upsample_add_bwd against torch autograd through the reference's own expressions
    (F.interpolate(x, 1/s), F.interpolate(flow, 1/s) / s, cat — IFNet.py:84-93 / :82-90; F.interpolate(head, s) * s + prev — :115-119)."""
    import torch.nn.functional as F
    from opticalflowscivis_b200 import _C, ops, train
    dev = _dev()
    torch.manual_seed(nd * 10 + scale)
    mode = "bilinear" if nd == 2 else "trilinear"
    n, nf = 2, 2 * nd
    img0, img1 = torch.rand((n, 1) + sp, device=dev), torch.rand((n, 1) + sp, device=dev)
    leaves = [torch.randn((n, c) + sp, device=dev, requires_grad=True) for c in (1, 1, 1, nf)]      # w0, w1, mask, flow
    # pack: ours
    xin = train._PackInputFn.apply(img0, img1, *leaves, scale)
    g = torch.randn(xin.shape, device=dev).bfloat16()
    g[..., 5 + nf:] = 0
    xin.backward(g)
    mine = [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    # pack: reference expression
    x = torch.cat((img0, img1, leaves[0], leaves[1], leaves[2]), 1)
    fl = leaves[3]
    if scale != 1:
        x = F.interpolate(x, scale_factor=1. / scale, mode=mode, align_corners=False)
    fl = F.interpolate(fl, scale_factor=1. / scale, mode=mode, align_corners=False) * 1. / scale
    ref_in = torch.cat((x, fl), 1)
    gref = g[..., :5 + nf].float()
    gref = gref[:, 0].permute(0, 3, 1, 2) if nd == 2 else gref.permute(0, 4, 1, 2, 3)
    ref_in.backward(gref)
    got_fwd = xin[..., :5 + nf].float()
    got_fwd = got_fwd[:, 0].permute(0, 3, 1, 2) if nd == 2 else got_fwd.permute(0, 4, 1, 2, 3)
    assert float((got_fwd - ref_in.detach()).abs().max()) <= 2e-2          # bf16 storage of the packed input
    for a, t in zip(mine, leaves):
        assert float((a - t.grad).abs().max()) <= 1e-5 * max(1.0, float(t.grad.abs().max()))
    # head up-sampling + accumulate
    hsp = tuple(v // scale for v in sp)
    head = torch.randn((n, 1 if nd == 2 else hsp[0]) + (hsp if nd == 2 else hsp[1:]) + (8,), device=dev, requires_grad=True)
    fprev, mprev = torch.randn((n, nf) + sp, device=dev, requires_grad=True), torch.randn((n, 1) + sp, device=dev, requires_grad=True)
    flow, mask = train._HeadUpFn.apply(head, fprev, mprev, scale, nd, sp)
    gf, gm = torch.randn_like(flow), torch.randn_like(mask)
    torch.autograd.backward((flow, mask), (gf, gm))
    gh, gfp, gmp = head.grad.clone(), fprev.grad.clone(), mprev.grad.clone()
    head.grad = fprev.grad = mprev.grad = None
    hn = head[:, 0].permute(0, 3, 1, 2) if nd == 2 else head.permute(0, 4, 1, 2, 3)
    flow_r = fprev + F.interpolate(hn[:, :nf], scale_factor=scale, mode=mode, align_corners=False, recompute_scale_factor=False) * scale
    mask_r = mprev + F.interpolate(hn[:, nf:nf + 1], scale_factor=scale, mode=mode, align_corners=False, recompute_scale_factor=False)
    assert float((flow - flow_r).abs().max()) <= 1e-5 * max(1.0, float(flow_r.abs().max()))
    torch.autograd.backward((flow_r, mask_r), (gf, gm))
    assert float((gh[..., :nf + 1] - head.grad[..., :nf + 1]).abs().max()) <= 2e-5 * max(1.0, float(head.grad.abs().max()))
    assert float(gh[..., nf + 1:].abs().max()) == 0.0
    assert torch.equal(gfp, fprev.grad) and torch.equal(gmp, mprev.grad)


@pytest.mark.parametrize("nd,c", [(3, 64), (2, 96)])
def test_train_block_batched_repack_equals_per_layer_pack(nd, c):
    """After the first pass created the packed bf16 operand forms, every refresh re-packs them IN PLACE with one launch
    (ofsv_conv_pack_weights_batched); the result must equal ofsv_conv_pack_weights of the refreshed tap forms, layer by layer."""
    from opticalflowscivis_b200 import ifnet, ops, train
    torch.manual_seed(9)
    dev = _dev()
    cin = 5 + 2 * nd
    blk = ifnet.IFBlock(nd, cin, c=c).to(dev)
    tb = train._TrainBlock(blk)
    x = torch.randn((1, cin) + (32,) * nd, device=dev, requires_grad=True)
    head = train._BlockFn.apply(x, tb, False, *blk.parameters())
    head.backward(torch.randn_like(head))                      # every forward and input-gradient layer has run: packed forms exist
    ptrs = {(i, k): v.data_ptr() for i, lay in enumerate(tb.fwd + tb.dgrad) for k, v in lay._packed.items()}
    assert len(ptrs) >= 24
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    tb.refresh()
    assert tb._pack_table is not None and tb._pack_table[1] == len(ptrs)
    for i, lay in enumerate(tb.fwd + tb.dgrad):
        for layout, w in lay._packed.items():
            assert w.data_ptr() == ptrs[(i, layout)]            # in place (CUDA graphs keep pointing at it)
            assert torch.equal(w, ops.conv_pack_weights(lay._structure_desc(), lay.w_simt, layout)), (i, layout)


@pytest.mark.parametrize("nd", [2, 3])
def test_conv_tc_paired_tiles_equal_single_tiles(nd):
    """ofsv_conv_tc with two output tiles per CTA sharing every weight tile (ofsv_set_tuning('tc_pair', 1)) computes the same MMAs in
    the same order per tile: bit-identical to the one-tile form on every layer family, odd tile counts and ragged edges included."""
    from opticalflowscivis_b200 import _C, ifnet, ops
    torch.manual_seed(21 + nd)
    dev = _dev()
    blk = ifnet.IFBlock(nd, 5 + 2 * nd, c=64).to(dev)
    L = blk.layers()
    n = 3
    for li, s in ((0, 24), (1, 12), (2, 8), (10, 6), (11, 12)):
        lay = L[li]
        in_sp = ((1,) if nd == 2 else ()) + (s,) * nd
        d, osp = lay.desc(n, in_sp, _C.BF16, has_residual=False)
        x = torch.randn((n,) + (((1,) + (s,) * 2) if nd == 2 else (s,) * 3) + (lay.cin_s,), device=dev).bfloat16()
        outs = []
        for pair in (0, 1):
            ops.set_tuning("tc_pair", pair)
            y = torch.full((n,) + tuple(osp) + (lay.cout_s,), float("nan"), device=dev, dtype=torch.float32 if lay.out_f32 else torch.bfloat16)
            ops.conv(d, x, lay.w_tc, lay.bias, lay.prelu, None, y, "tc")
            outs.append(y)
        ops.set_tuning("tc_pair", -1)
        assert torch.isfinite(outs[1].float()).all(), li
        assert torch.equal(outs[0], outs[1]), li
