"""Bring-up probe for the tcgen05 conv engine (run on the GPU box under `timeout`): one layer per K-chunk width /
layer type against the CPU tap evaluator.  Prints per-case max error; exits non-zero on the first failure."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from opticalflowscivis_b200 import _C, ifnet, ops  # noqa: E402
from tap_eval import run_layer  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    bad = 0
    engine = "tc"
    args = sys.argv[1:]
    if args and args[0] in ("tc", "halo"):
        engine, args = args[0], args[1:]
    only = args or None
    for nd in (3, 2):
        for c in (64, 32, 128):
            cin = 5 + 2 * nd
            blk = ifnet.IFBlock(nd, cin, c)
            for p in blk.parameters():
                if p.dim() == 1 and float(p.data.std()) == 0:
                    p.data.uniform_(0.05, 0.5)
            blk_dev = ifnet.IFBlock(nd, cin, c).to(dev)
            blk_dev.load_state_dict(blk.state_dict())
            Lc, Ld = blk.layers(), blk_dev.layers()
            sp = (1, 24, 40) if nd == 2 else ((12, 8, 16) if engine == "tc" else (10, 24, 20))
            for li in ((2, 3, 10, 11) if engine == "halo" else (2, 3, 1, 0, 10, 11)):
                tag = f"nd{nd}_c{c}_L{li}"
                if only and tag not in only:
                    continue
                lc, ld = Lc[li], Ld[li]
                if engine == "halo" and li == 11:
                    lc, ld = blk._heads_shuffle, blk_dev._heads_shuffle
                xin = (torch.randn((2,) + sp + (lc.cin_s,)) * 0.5).bfloat16().float()
                d, osp = lc.desc(2, sp, _C.BF16)
                res = (torch.randn((2,) + osp + (lc.cout_s,)) * 0.5).bfloat16().float() if lc.residual else None
                ref = run_layer(lc, xin, res)
                y = torch.full((2,) + osp + (lc.cout_s,), float("nan"), device=dev,
                               dtype=torch.float32 if lc.out_f32 else torch.bfloat16)
                try:
                    ops.conv(d, xin.to(dev).bfloat16(), ld.w_tc, ld.bias, ld.prelu,
                             None if res is None else res.to(dev).bfloat16(), y, engine)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    print(f"{tag}: EXCEPTION {e}", flush=True)
                    sys.exit(2)
                got = y.float().cpu().view(ref.shape)
                err = float((got - ref).abs().max())
                nan = int(torch.isnan(got).sum())
                scale = float(ref.abs().max())
                ok = nan == 0 and err <= 3e-2 * max(1.0, scale)
                print(f"{tag}: kc={lc.kc} N={lc.cout_w} taps={lc.ntaps}x{lc.nphase} stride={lc.in_stride} "
                      f"max|err|={err:.3e} (ref max {scale:.2f}) nan={nan} {'ok' if ok else 'FAIL'}", flush=True)
                bad += not ok
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
