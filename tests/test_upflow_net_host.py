"""CPU test of the UPFlow network's host wiring (no kernels are launched): opticalflowscivis_b200/upflow/net.py
with every libofsv call replaced by the fp32 torch evaluator of the same contract (tap-form convolutions through tests/tap_eval.py,
operators through oracle/) must reproduce the record of the REFERENCE's `UPFlow_net.forward_2_frame_v3`
(tests/golden/upflow_net.npz, written by tests/golden/make_upflow_net_golden.py from the imported reference)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from opticalflowscivis_b200.upflow import net as unet
from oracle import ops_ref, upflow_ref as ur
from tap_eval import run_layer

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_tap_layer_eval(lay, x, n, in_sp):
    d, osp = lay.desc(n, in_sp, 0, has_residual=False)
    return run_layer(lay, x.float()), osp


def _from_cl_f32(y, c, nd):
    return y[:, 0, :, :, :c].permute(0, 3, 1, 2).float().contiguous()


def _to_cl_f32(x, nd, cs):
    if isinstance(x, (list, tuple)):
        x = torch.cat(list(x), 1)
    n, c = x.shape[:2]
    out = torch.zeros((n, 1) + tuple(x.shape[2:]) + (cs,))
    out[:, 0, :, :, :c] = x.permute(0, 2, 3, 1)
    return out


@pytest.fixture
def cpu_engine(monkeypatch):
    monkeypatch.setattr(unet, "run_tap_layer", _run_tap_layer_eval)
    monkeypatch.setattr(unet, "_to_cl", _to_cl_f32)
    monkeypatch.setattr(unet, "_from_cl", _from_cl_f32)
    monkeypatch.setattr(unet.ops, "_cuda_f32", lambda t, name: t.float())
    monkeypatch.setattr(unet.ops, "upsample_flow_ac", lambda f, h, w, if_rate=True: ops_ref.upsample2d_flow_as_ref(f, h, w, if_rate))
    monkeypatch.setattr(unet.ops, "warping_no_div", ops_ref.warping_layer_no_div_ref)
    monkeypatch.setattr(unet.ops, "torch_warp", ur.torch_warp_ref)
    monkeypatch.setattr(unet.ops, "corr81_fwd", lambda a, b, leaky_slope=None: F.leaky_relu(ops_ref.corr81_ref(a, b), leaky_slope))
    monkeypatch.setattr(unet.ops, "feature_norm_pair",
                        lambda a, b, flow=None: tuple(ur.normalize_features_ref(a, b if flow is None else ops_ref.warping_layer_no_div_ref(b, flow))))


@pytest.mark.parametrize("tag,sgu", [("train", False), ("sgu", True)])
def test_upflow_net_wiring_reproduces_reference_record(tag, sgu, cpu_engine):
    gold = np.load(os.path.join(G, "upflow_net.npz"))
    net = unet.UPFlowNet(if_sgu_upsample=sgu)
    sd = net.state_dict()
    assert sorted(sd) == list(gold[f"{tag}_names"])                                  # the reference's state_dict keys ...
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == list(gold[f"{tag}_shapes"])   # ... and shapes
    net.load_state_dict(ur.deterministic_state({k: tuple(v.shape) for k, v in sd.items()}))
    im1, im2 = ur.smooth_pair(1, 128, 192)
    ff, fb, flows = net.forward_2_frame_v3(im1, im2)
    # Free-running: exact on the coarse levels.  From the level at which the warped features first straddle WarpingLayer_no_div's
    # `grid_sample(ones) >= 1` mask (pwc_modules.py:203-206) the comparison is chaotic BY CONSTRUCTION OF THE REFERENCE: the test
    # is decided by the last bit of the sum of four bilinear weights, so a 1e-8 difference in the incoming flow zeroes or keeps
    # whole feature vectors.  Those levels are therefore checked teacher-forced below (recorded flow of the level before as input).
    for i in (4, 3, 2):
        # (with the self-guided up-sampling a masked warp sits INSIDE every level above the coarsest: free-running, a changed fp32
        # summation order already flips mask pixels there; those levels are pinned by the teacher-forced check)
        assert np.abs(flows[i][0].numpy() - gold[f"{tag}_lvl{i}_f"]).max() <= 1e-5 + (5e-3 if sgu else 0), i
        assert np.abs(flows[i][1].numpy() - gold[f"{tag}_lvl{i}_b"]).max() <= 1e-5 + (5e-3 if sgu else 0), i
    for i in (1, 0):
        ref = gold[f"{tag}_lvl{i}_f"]
        assert np.abs(flows[i][0].numpy() - ref).mean() <= 0.05 * np.abs(ref).mean(), i
    errs = teacher_forced_level_errors(net, gold, tag, im1, im2)
    assert max(errs) <= 2e-5, errs
    if not sgu:
        of, ob = unet.occ_check(torch.from_numpy(gold[f"{tag}_lvl0_f"]), torch.from_numpy(gold[f"{tag}_lvl0_b"]))
        rf, rb = ur.occ_check_ref(torch.from_numpy(gold[f"{tag}_lvl0_f"]), torch.from_numpy(gold[f"{tag}_lvl0_b"]))
        assert torch.equal(of, rf) and torch.equal(ob, rb)


def teacher_forced_level_errors(net, gold, tag, im1, im2):
    """Max-abs error of every pyramid level's output flow when the level is fed the RECORDED flow of the level before (zeros at the
    coarsest) — the warps, and with them the validity masks, are then the reference's own, and the level's arithmetic is compared
    like for like.  Shared by the CPU wiring test and the GPU parity test (tests/test_gpu_train.py)."""
    b = im1.shape[0]
    both = torch.cat((im1, im2), 0)
    swap = lambda t: torch.cat((t[b:], t[:b]), 0)                               # noqa: E731
    pyramid = net.feature_pyramid_extractor(both)
    errs = []
    for level in range(5):
        feat = pyramid[level]
        if level == 0:
            flow = torch.zeros((2 * b, 2) + tuple(feat.shape[2:]), device=both.device)
        else:
            flow = torch.cat((torch.from_numpy(gold[f"{tag}_lvl{5 - level}_f"]), torch.from_numpy(gold[f"{tag}_lvl{5 - level}_b"])), 0).to(both.device)
        flow_up, res = net.decode_level_res(level, flow, feat, net._conv1x1(level, feat), swap)
        out = (flow_up + res).cpu()
        ref = torch.cat((torch.from_numpy(gold[f"{tag}_lvl{4 - level}_f"]), torch.from_numpy(gold[f"{tag}_lvl{4 - level}_b"])), 0)
        errs.append(float((out - ref).abs().max()))
    # full resolution from the recorded output-level flow (every 4th pixel was recorded)
    flow = torch.cat((torch.from_numpy(gold[f"{tag}_lvl0_f"]), torch.from_numpy(gold[f"{tag}_lvl0_b"])), 0).to(both.device)
    out = net.upsample_output(flow, both, swap).cpu()[:, :, ::4, ::4]
    ref = torch.cat((torch.from_numpy(gold[f"{tag}_flow_f_sub4"]), torch.from_numpy(gold[f"{tag}_flow_b_sub4"])), 0)
    errs.append(float((out - ref).abs().max()))
    return errs


def test_upflow_net_rejects_unsupported_configurations():
    with pytest.raises(NotImplementedError):
        unet.UPFlowNet(norm_moments_across_channels=True)
    with pytest.raises(NotImplementedError):
        unet.UPFlowNet().forward()
