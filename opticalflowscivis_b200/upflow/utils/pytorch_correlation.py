"""UPFlow/utils/pytorch_correlation.py:10-50 — `Corr_pyTorch`, the cost-volume module UPFlow selects with `if_use_cor_pytorch`
(UPFlow/model/upflow.py:359-361, 643-645).  Same constructor and call surface; the arithmetic is the correlation kernel behind
`CorrelationFunction` (ofsv_corr81_{fwd,bwd}_f32) instead of the reference's unfold / (B,81,C,HW) product / mean chain, which it
equals to summation order (tests/golden: 2e-6).  Differentiable in both inputs like the reference module."""
from __future__ import annotations

import torch.nn as nn

from ..correlation import CorrelationFunction


class Corr_pyTorch(nn.Module):   # noqa: N801  (reference class name)
    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1):
        assert pad_size == max_displacement                      # the reference's own asserts (pytorch_correlation.py:17-18)
        assert stride1 == stride2 == 1
        super().__init__()
        self.pad_size, self.kernel_size, self.stride1, self.stride2 = pad_size, kernel_size, stride1, stride2
        self.max_hdisp = max_displacement
        self.corr_multiply = corr_multiply

    def forward(self, in1, in2):
        return CorrelationFunction.apply(in1, in2, self.pad_size, self.kernel_size, self.max_hdisp, self.stride1, self.stride2,
                                         self.corr_multiply)
