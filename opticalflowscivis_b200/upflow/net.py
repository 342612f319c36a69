"""UPFlow's flow network on libofsv: `UPFlow_net.forward_2_frame_v3` — UPFlow/model/upflow.py:580-620 — with
`decode_level_res` (:621-665), `network_tools.normalize_features` (:95-138), `sgu_model` (:21-93) and the forward-backward occlusion
check (UPFlow/utils/tools.py:592-630).  Inference path (no autograd); the training-step operators live in ops.py / upflow_bwd.cu.

Parameter containers with the reference's `state_dict()` keys (`feature_pyramid_extractor.convs.0.0.0.weight` …
`context_networks.convs.6.0.bias`, `conv_1x1.4.0.weight`, `sgi_model.dense_estimator_mask.conv1.0.weight` …).  What runs where:
  * every convolution (feature pyramid, 1x1 reductions, the dense flow estimator, the dilated context network, the sgu nets) is a
    tap-form layer on the tcgen05 engines of the IFBlocks — bf16 operands, fp32 accumulate, LeakyReLU(0.1) as a PReLU epilogue with
    a constant slope, dilation as tap offsets, the dense `torch.cat([conv(x), x])` as a channels-last concatenation whose padding
    stays at the end (so the reference's input-channel order is the physical one); the two 2-channel flow heads write fp32;
  * the producer of the cost-volume inputs is ONE launch per level for BOTH directions (ofsv_feature_norm_pair_f32: feature warp
    by WarpingLayer_no_div + per-plane normalisation of both tensors; the directions are stacked along the batch), followed by the
    81-channel correlation with the LeakyReLU fused (ofsv_corr81_fwd_f32);
  * flow resizes: ofsv_upsample_flow_ac_f32; occlusion check / sgu interpolation: ofsv_torch_warp_f32.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ifnet import _ConvParams, _Layer, _rup
from ..refine import _from_cl, _to_cl, run_tap_layer

NUM_CHS = [3, 16, 32, 64, 96, 128, 196]        # upflow.py:337


def _conv(cin, cout, kernel_size=3, stride=1, dilation=1, isReLU=True):
    """pwc_modules.conv (:10-50, plain branch): Conv2d(padding = (k-1)*dilation//2) [+ LeakyReLU(0.1)]; initialize_msra (:53-70)."""
    m = _ConvParams(2, cin, cout, kernel_size, stride, ((kernel_size - 1) * dilation) // 2)
    m.dilation, m.leaky = dilation, bool(isReLU)
    nn.init.kaiming_normal_(m.weight)
    nn.init.constant_(m.bias, 0)
    return nn.Sequential(m)


def _pad64(c):
    """Physical width of a c-channel activation: wide ones are padded to a multiple of 64 so that the NEXT layer's K chunks are 64
    channels (ofsv_conv_tc picks KC = 64 / 32 / 16 from Cin_s; with 16-channel chunks a 592-channel layer makes 37 tiny K steps per
    tap: 0.53 ms for the first context layer of the finest level, `profiles/r03p_ncu_full_upflow_convs.csv`)."""
    return _rup(c, 64) if c >= 96 else _rup(c, 16)


def _layers(m, in_map=None, cin_phys=None, out_f32=False, cout_phys=None):
    """Tap-form layer(s) of one pwc_modules.conv: dilation as tap offsets, LeakyReLU(0.1) as a constant PReLU slope, the input
    channels scattered to their PHYSICAL positions (in_map; identity when only cin_phys is given: padding at the end), the output
    zero-padded to cout_phys channels and split into <= 128-channel launches (ofsv_conv_tc's limit; FeatureExtractor's last level
    has 196 -> 128 + 128)."""
    k, d = m.k, m.dilation
    w = m.weight.detach().float()
    cin = w.shape[1]
    if cin_phys is not None:
        wp = torch.zeros((w.shape[0], cin_phys) + tuple(w.shape[2:]), device=w.device)
        wp[:, in_map if in_map is not None else list(range(cin))] = w
        w = wp
    bias = m.bias.detach().float()
    if cout_phys is not None and cout_phys > m.cout and not out_f32:
        w = torch.cat((w, torch.zeros((cout_phys - m.cout,) + tuple(w.shape[1:]), device=w.device)), 0)
        bias = torch.cat((bias, torch.zeros(cout_phys - m.cout, device=w.device)))
    cout = w.shape[0]
    taps = [(0, (ky - (k - 1) // 2) * d, (kx - (k - 1) // 2) * d) for ky in range(k) for kx in range(k)]
    out = []
    for lo in range(0, cout, 128):
        hi = min(cout, lo + 128)
        w_tap = torch.stack([w[lo:hi, :, ky, kx].t() for ky in range(k) for kx in range(k)])        # [T][Cin][Cout]
        slope = torch.full((hi - lo,), 0.1, device=w.device) if m.leaky else None
        lay = _Layer(2, m.stride, 1, 1, taps, w_tap, bias[lo:hi], slope, 8 if out_f32 else _rup(hi - lo, 16), out_f32=out_f32)
        out.append(lay)
    return out


class _Net(nn.Module):
    """Caches the tap-form layers of its convs per parameter version."""

    def _packed_layers(self, build):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_pk", None) != key:
            self._pl, self._pk = build(), key
        return self._pl


def _run(lays, x, n, sp):
    ys = [run_tap_layer(lay, x, n, sp) for lay in lays]
    return (ys[0][0] if len(ys) == 1 else torch.cat([y for y, _ in ys], -1)), ys[0][1]


class FeatureExtractor(_Net):
    """pwc_modules.py:124-143: six levels of conv(stride 2) + conv; returns the pyramid coarsest first, fp32 (B,C,h,w)."""

    def __init__(self, num_chs=NUM_CHS):
        super().__init__()
        self.num_chs = num_chs
        self.convs = nn.ModuleList(nn.Sequential(_conv(a, b, stride=2), _conv(b, b)) for a, b in zip(num_chs[:-1], num_chs[1:]))

    @torch.no_grad()
    def forward(self, x):
        def build():
            out, cin_phys = [], 16
            for st in self.convs:
                cp = _pad64(st[0][0].cout)
                out.append([_layers(st[0][0], cin_phys=cin_phys, cout_phys=cp), _layers(st[1][0], cin_phys=cp, cout_phys=cp)])
                cin_phys = cp
            return out
        L = self._packed_layers(build)
        n, sp = x.shape[0], (1,) + tuple(x.shape[2:])
        a = _to_cl(x, 2, 16)
        out = []
        for l, (la, lb) in enumerate(L):
            a, sp = _run(la, a, n, sp)
            a, sp = _run(lb, a, n, sp)
            out.append(_from_cl(a, self.num_chs[l + 1], 2))
        return out[::-1]


class FlowEstimatorDense(_Net):
    """pwc_modules.FlowEstimatorDense_v2 (:250-286) and sgu_model's FlowEstimatorDense_temp (upflow.py:25-60): five dense 3x3 convs
    (x_{i+1} = cat([conv_i(x_i), x_i])) and a linear head.  Returns (x5 fp32 (B, N, h, w), head fp32 (B, ch_out, h, w))."""

    def __init__(self, ch_in, f_channels=(128, 128, 96, 64, 32), out_channel=2):
        super().__init__()
        self.ch_in, self.f_channels, self.ch_out = ch_in, tuple(f_channels), out_channel
        n = ch_in
        for i, f in enumerate(f_channels):
            setattr(self, f"conv{i + 1}", _conv(n, f))
            n += f
        self.n_channels = n
        self.conv_last = _conv(n, out_channel, isReLU=False)

    def in_phys(self):
        """Physical width of the input tensor `run` expects."""
        return _rup(self.ch_in, 16)

    def _build(self):
        # physical channels-last layout of x_i: [conv_i out | ... | conv_1 out | x (ch_in, zero-padded)], every piece padded on its own
        # (to a multiple of 64 for the wide main estimator: the running totals 128, 256, 384, 512, 576, 640 then all give KC = 64) and
        # the reference's logical input-channel order mapped onto it
        phys = self.in_phys()
        maps = [list(range(self.ch_in))]                       # logical input channel -> physical position, per layer
        pieces = [(self.ch_in, phys)]                          # (logical, physical) widths, newest first
        L = []
        for i, f in enumerate(self.f_channels):
            m = getattr(self, f"conv{i + 1}")[0]
            fp = _rup(f, 16)
            if self.ch_in >= 96 and phys % 64 == 0:            # keep the running total a multiple of 64 (96 -> 128, the last 32 -> 64)
                fp = _rup(f, 64) if (f >= 96 or i == len(self.f_channels) - 1) else fp
            L.append(_layers(m, in_map=maps[-1], cin_phys=phys, cout_phys=fp))
            pieces.insert(0, (f, fp))
            phys += fp
            mp, off = [], 0
            for lg, ph in pieces:
                mp += list(range(off, off + lg))
                off += ph
            maps.append(mp)
        L.append(_layers(self.conv_last[0], in_map=maps[-1], cin_phys=phys, out_f32=True))
        self._x5_map = maps[-1]
        return L

    @torch.no_grad()
    def run(self, x_cl, n, sp):
        """x_cl: channels-last bf16 [B][1][h][w][rup(ch_in,16)].  Returns (x5 channels-last bf16 physical, x5 map, head fp32)."""
        L = self._packed_layers(self._build)
        x = x_cl
        for lays in L[:-1]:
            y, _ = _run(lays, x, n, sp)
            x = torch.cat((y, x), -1)
        head, _ = _run(L[-1], x, n, sp)
        return x, self._x5_map, _from_cl(head, self.ch_out, 2)


class ContextNetwork(_Net):
    """pwc_modules.ContextNetwork_v2_ (:396-412): 3x3 convs with dilations 1, 2, 4, 8, 16, 1 and a linear 2-channel head."""

    def __init__(self, ch_in, f_channels=(128, 128, 128, 96, 64, 32, 2)):
        super().__init__()
        dil = (1, 2, 4, 8, 16, 1)
        chans = (ch_in,) + tuple(f_channels)
        self.convs = nn.Sequential(*[_conv(chans[i], chans[i + 1], 3, 1, dil[i]) for i in range(6)],
                                   _conv(chans[6], chans[7], isReLU=False))

    @torch.no_grad()
    def run(self, x_cl, in_map, n, sp):
        def build():
            cp = [_pad64(self.convs[i][0].cout) for i in range(6)]
            L = [_layers(self.convs[0][0], in_map=in_map, cin_phys=x_cl.shape[-1], cout_phys=cp[0])]
            L += [_layers(self.convs[i][0], cin_phys=cp[i - 1], cout_phys=cp[i]) for i in range(1, 6)]
            L.append(_layers(self.convs[6][0], cin_phys=cp[5], out_f32=True))
            return L
        L = self._packed_layers(build)
        x = x_cl
        for lays in L:
            x, _ = _run(lays, x, n, sp)
        return _from_cl(x, 2, 2)


class SguModel(_Net):
    """network_tools.sgu_model (upflow.py:21-93): self-guided up-sampling — an interpolation flow and mask from a small dense
    estimator on (feature_1, warped feature_2); flow_up = torch_warp(flow, inter_flow) * (1 - mask) + flow * mask."""

    def __init__(self):
        super().__init__()
        self.dense_estimator_mask = FlowEstimatorDense(64, f_channels=(32, 32, 32, 16, 8), out_channel=3)
        self.upsample_output_conv = nn.Sequential(_conv(3, 16), _conv(16, 16, stride=2), _conv(16, 32), _conv(32, 32, stride=2))

    @torch.no_grad()
    def output_conv(self, x):
        L = self._packed_layers(lambda: [_layers(s[0]) for s in self.upsample_output_conv])
        n, sp = x.shape[0], (1,) + tuple(x.shape[2:])
        a = _to_cl(x, 2, 16)
        for lays in L:
            a, sp = _run(lays, a, n, sp)
        return _from_cl(a, 32, 2)

    @torch.no_grad()
    def forward(self, flow_init, feature_1, feature_2, output_level_flow=None):
        n, _, h, w = flow_init.shape
        hf, wf = feature_1.shape[2:]
        if (h, w) != (hf, wf):
            flow_init = ops.upsample_flow_ac(flow_init, hf, wf)
        f2w = ops.warping_no_div(feature_2, flow_init)
        x_cl = _to_cl([feature_1, f2w], 2, _rup(feature_1.shape[1] + f2w.shape[1], 16))
        _, _, x_out = self.dense_estimator_mask.run(x_cl, n, (1, hf, wf))
        inter_flow, inter_mask = x_out[:, :2].contiguous(), torch.sigmoid(x_out[:, 2:3])
        if output_level_flow is not None:
            H, W = output_level_flow.shape[2:]
            inter_flow = ops.upsample_flow_ac(inter_flow, H, W)
            inter_mask = ops.upsample_flow_ac(torch.cat((inter_mask, inter_mask), 1).contiguous(), H, W, if_rate=False)[:, :1]
            flow_init = output_level_flow
        flow_up = ops.torch_warp(flow_init, inter_flow) * (1 - inter_mask) + flow_init * inter_mask
        return flow_init, flow_up, inter_flow, inter_mask


class UPFlowNet(nn.Module):
    """UPFlow_net (upflow.py:290-367), inference: `forward_2_frame_v3(x1_raw, x2_raw)` -> (flow_f, flow_b, flows[::-1])."""

    def __init__(self, if_norm_before_cost_volume=True, norm_moments_across_channels=False, norm_moments_across_images=False,
                 if_sgu_upsample=False):
        super().__init__()
        if if_norm_before_cost_volume and (norm_moments_across_channels or norm_moments_across_images):
            raise NotImplementedError("normalize_features: only the per-(sample, channel) moments of scripts/simple_train.py:321-329 / "
                                      "test.py:116-118 are provided (moments_across_channels = moments_across_images = False)")
        self.if_norm, self.if_sgu = bool(if_norm_before_cost_volume), bool(if_sgu_upsample)
        self.output_level = 4
        self.feature_pyramid_extractor = FeatureExtractor()
        self.flow_estimators = FlowEstimatorDense(81 + 32 + 2)
        self.context_networks = ContextNetwork(self.flow_estimators.n_channels + 2)
        self.conv_1x1 = nn.ModuleList(_conv(c, 32, kernel_size=1) for c in (196, 128, 96, 64, 32))
        self.sgi_model = SguModel() if self.if_sgu else None

    def _conv1x1(self, l, x):
        net = self.conv_1x1[l]
        key = (net[0].weight.data_ptr(), net[0].weight._version, net[0].bias._version)
        if getattr(net, "_pk", None) != key:
            net._pl, net._pk = _layers(net[0], cin_phys=_pad64(net[0].cin)), key
        n, sp = x.shape[0], (1,) + tuple(x.shape[2:])
        y, _ = _run(net._pl, _to_cl(x, 2, _pad64(x.shape[1])), n, sp)
        return _from_cl(y, 32, 2)

    @torch.no_grad()
    def decode_level_res(self, level, flow, feat, feat_1x1, swap):
        """upflow.py:621-665 for BOTH directions at once: batch = [forward pairs ; backward pairs], `swap` exchanges the halves."""
        h, w = feat.shape[2:]
        n = feat.shape[0]
        flow_up = ops.upsample_flow_ac(flow, h, w)
        other, other_1x1 = swap(feat), swap(feat_1x1)
        if level > 0 and self.if_sgu:
            _, flow_up, _, _ = self.sgi_model(flow_up, feat_1x1, other_1x1)
        if self.if_norm:
            f1, f2w = ops.feature_norm_pair(feat, other, flow_up if level > 0 else None)
        else:
            f1, f2w = feat, (ops.warping_no_div(other, flow_up) if level > 0 else other)
        corr = ops.corr81_fwd(f1, f2w, leaky_slope=0.1)
        x_cl = _to_cl([corr, feat_1x1, flow_up], 2, _rup(corr.shape[1] + feat_1x1.shape[1] + 2, 16))    # upflow.py:657, one launch
        x5, x5_map, flow_res = self.flow_estimators.run(x_cl, n, (1, h, w))
        flow_ = flow_up + flow_res
        ctx_in = torch.cat((x5, _to_cl(flow_, 2, 64 if x5.shape[-1] % 64 == 0 else 16)), -1)     # 576 + 64 = 640 channels: KC = 64
        ctx_map = x5_map + [x5.shape[-1], x5.shape[-1] + 1]
        flow_fine = self.context_networks.run(ctx_in, ctx_map, n, (1, h, w))
        return flow_up, flow_res + flow_fine

    def enable_cuda_graphs(self, on=True):
        """Replay `forward_2_frame_v3` from a CUDA graph captured per input shape: a call is ~520 libofsv launches plus the torch
        layout copies, more host time than GPU time at 256 x 832.  Same numerics; the inputs are copied into graph-owned buffers
        and the returned tensors are the graph's outputs, overwritten by the next call with the same shape (clone what you keep).
        Graphs are dropped when a parameter changes."""
        self._graphs = {} if on else None
        return self

    @torch.no_grad()
    def forward_2_frame_v3(self, x1_raw, x2_raw, if_loss=False):
        graphs = getattr(self, "_graphs", None)
        if graphs is None:
            return self._forward_2_frame_v3(x1_raw, x2_raw)
        x1_raw, x2_raw = ops._cuda_f32(x1_raw, "x1_raw"), ops._cuda_f32(x2_raw, "x2_raw")
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if graphs.get("sig") != sig:
            graphs.clear()
            graphs["sig"] = sig
        key = (tuple(x1_raw.shape), x1_raw.device.index)
        ent = graphs.get(key)
        if ent is None:
            a, b = x1_raw.clone(), x2_raw.clone()
            cur = torch.cuda.current_stream(a.device)
            side = torch.cuda.Stream(device=a.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                   # warm-up off the capture: packed weights, engine choice, kernel attributes
                for _ in range(2):
                    self._forward_2_frame_v3(a, b)
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward_2_frame_v3(a, b)
            ent = graphs[key] = (g, a, b, out)
        g, a, b, out = ent
        a.copy_(x1_raw)
        b.copy_(x2_raw)
        g.replay()
        return out

    @torch.no_grad()
    def _forward_2_frame_v3(self, x1_raw, x2_raw):
        x1_raw, x2_raw = ops._cuda_f32(x1_raw, "x1_raw"), ops._cuda_f32(x2_raw, "x2_raw")
        if x1_raw.dim() != 4 or x1_raw.shape[1] != 3 or x2_raw.shape != x1_raw.shape:
            raise ValueError(f"UPFlowNet: expected two (B,3,H,W) images, got {tuple(x1_raw.shape)} / {tuple(x2_raw.shape)}")
        if x1_raw.shape[2] % 64 or x1_raw.shape[3] % 64:
            raise NotImplementedError("UPFlowNet: H and W must be multiples of 64 (six stride-2 levels)")
        b = x1_raw.shape[0]
        swap = lambda t: torch.cat((t[b:], t[:b]), 0)                               # noqa: E731
        both = torch.cat((x1_raw, x2_raw), 0)                                      # [x1 ; x2]: the forward / backward directions
        pyramid = self.feature_pyramid_extractor(both) + [both]
        flow = torch.zeros((2 * b, 2) + tuple(pyramid[0].shape[2:]), device=both.device)
        flows = []
        for level in range(self.output_level + 1):
            feat = pyramid[level]
            feat_1x1 = self._conv1x1(level, feat)
            flow, flow_res = self.decode_level_res(level, flow, feat, feat_1x1, swap)
            flow = flow + flow_res
            flows.append([flow[:b], flow[b:]])
        flow_out = self.upsample_output(flow, both, swap)
        return flow_out[:b], flow_out[b:], flows[::-1]

    @torch.no_grad()
    def upsample_output(self, flow, both, swap):
        """upflow.py:609-619: the output-level flow to full resolution (bilinear, then the self-guided correction when enabled)."""
        H, W = both.shape[2:]
        flow_out = ops.upsample_flow_ac(flow, H, W)
        if self.if_sgu:
            f_1x1 = self.sgi_model.output_conv(both)
            _, flow_out, _, _ = self.sgi_model(flow, f_1x1, swap(f_1x1), output_level_flow=flow_out)
        return flow_out

    def forward(self, *a, **k):
        raise NotImplementedError("UPFlowNet provides the inference path `forward_2_frame_v3`; the loss / training `forward` of "
                                  "upflow.py:428-578 is outside this tier")


def occ_check(flow_fw, flow_bw, alpha_1=0.1, alpha_2=0.5, scale=1, obj_out_all="obj"):
    """tools.occ_check_model(occ_type='for_back_check') — UPFlow/utils/tools.py:543-719; 1 = consistent, 0 = occluded.  The
    forward-backward check (:592-630) compares |flow + warp(opposite flow)| (sum of absolute components) with alpha_1 * (|fw| + |bw|)
    + alpha_2 / scale; with obj_out_all = 'obj' (UPFlow_net's setting, upflow.py:299) pixels whose flow leaves the frame (:683-710)
    are not counted as occluded (:713-719).  The two flow warps run on ofsv_torch_warp_f32."""
    flow_fw, flow_bw = ops._cuda_f32(flow_fw, "flow_fw"), ops._cuda_f32(flow_bw, "flow_bw")
    if obj_out_all not in ("obj", "all"):
        raise NotImplementedError("occ_check: obj_out_all must be 'obj' or 'all'")
    length = lambda x: x.abs().sum(1, keepdim=True)                                 # noqa: E731  (sum_abs_or_squar is forced True, :558)
    mag = length(flow_fw) + length(flow_bw)
    diff_fw = flow_fw + ops.torch_warp(flow_bw, flow_fw)
    diff_bw = flow_bw + ops.torch_warp(flow_fw, flow_bw)
    thresh = alpha_1 * mag + alpha_2 / scale
    occ = [(length(diff_fw) < thresh), (length(diff_bw) < thresh)]
    if obj_out_all == "all":
        return occ[0].float(), occ[1].float()
    out = []
    for o, fl in zip(occ, (flow_fw, flow_bw)):
        _, _, h, w = fl.shape
        px = torch.arange(w, device=fl.device).view(1, 1, 1, w).float() + fl[:, 0:1]
        py = torch.arange(h, device=fl.device).view(1, 1, h, 1).float() + fl[:, 1:2]
        inside = (px <= w - 1) & (px >= 0) & (py <= h - 1) & (py >= 0)
        out.append((o | ~inside).float())
    return out[0], out[1]
