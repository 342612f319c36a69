"""CPU evaluator of the library's tap-form convolution descriptor (test infrastructure).

Computes exactly what include/ofsv.h specifies for `ofsv_conv_desc` with torch indexing on CPU, so the weight packing
and the ConvTranspose phase decomposition of opticalflowscivis_b200/ifnet.py can be checked against torch's own
conv / conv_transpose without a GPU."""
import torch


def run_layer(lay, x, residual=None):
    """lay: ifnet._Layer ; x: [N][D][H][W][Cin_s] fp32 (D = 1 for 2-D) -> y [N][Dy][Hy][Wy][Cout_s]
    (always the plain logical layout, also for `out_s2d` layers)."""
    n, di, hi, wi, _ = x.shape
    if getattr(lay, "in_s2d", False):       # x is the shifted space-to-depth tensor: desc() wants the logical dims
        d, osp = lay.desc(n, ((di - 1) * 2 if lay.nd == 3 else 1, (hi - 1) * 2, (wi - 1) * 2), 0)
    else:
        d, osp = lay.desc(n, (di, hi, wi), 0)
    y = torch.zeros(n, osp[0], osp[1], osp[2], lay.cout_s)
    w = lay.w_simt.cpu()
    zs, ys, xs = torch.arange(d.Do), torch.arange(d.Ho), torch.arange(d.Wo)
    for ph in range(lay.nphase):
        acc = lay.bias.cpu().view(1, 1, 1, 1, -1).expand(n, d.Do, d.Ho, d.Wo, lay.cout_w).clone()
        for t in range(lay.ntaps):
            oz, oy, ox = lay.taps[ph * lay.ntaps + t]
            iz, iy, ix = zs * lay.in_stride + oz, ys * lay.in_stride + oy, xs * lay.in_stride + ox
            vz, vy, vx = (iz >= 0) & (iz < di), (iy >= 0) & (iy < hi), (ix >= 0) & (ix < wi)
            g = x[:, iz.clamp(0, di - 1)][:, :, iy.clamp(0, hi - 1)][:, :, :, ix.clamp(0, wi - 1)]
            valid = (vz.view(-1, 1, 1) & vy.view(1, -1, 1) & vx.view(1, 1, -1)).view(1, d.Do, d.Ho, d.Wo, 1)
            acc = acc + torch.einsum("ndhwc,co->ndhwo", g * valid, w[ph * lay.ntaps + t])
        if lay.prelu is not None:
            a = lay.prelu.cpu().view(1, 1, 1, 1, -1)
            acc = torch.where(acc > 0, acc, acc * a)
        if getattr(lay, "shuffle", 0):
            npar = lay.cout_w // lay.shuffle
            for q in range(npar):           # columns [output parity][8 channels] -> depth-to-space
                qz, qy, qx = ((q >> 2) & 1, (q >> 1) & 1, q & 1) if lay.nd == 3 else (0, (q >> 1) & 1, q & 1)
                y[:, qz::(2 if lay.nd == 3 else 1), qy::2, qx::2] = acc[..., q * 8:(q + 1) * 8]
            continue
        pz, py, px = (ph >> 2) & 1, (ph >> 1) & 1, ph & 1
        y[:, pz::lay.out_stride, py::lay.out_stride, px::lay.out_stride] = acc[..., : lay.cout_s]
    if residual is not None:
        y = y + residual
    return y
