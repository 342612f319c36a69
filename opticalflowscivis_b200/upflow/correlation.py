"""UPFlow/model/correlation_package/correlation.py:6-45 — `CorrelationFunction` — plus a `correlation_cuda`
shim with the un-vendored extension's `forward` / `backward` signatures (correlation.py:26-27,42-43)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import ops

_ONLY = (4, 1, 4, 1, 1, 1)   # the single configuration used by the reference (UPFlow/model/upflow.py:649,652)


def _check(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
    cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
    if cfg != _ONLY:
        raise NotImplementedError(f"correlation config {cfg} is never used by the reference; only {_ONLY} is implemented")


class correlation_cuda:   # noqa: N801  (module-like namespace mirroring the pybind extension)
    """`correlation_cuda.forward(input1, input2, rbot1, rbot2, output, pad, k, md, s1, s2, mult)`: the caller passes
    EMPTY tensors that the extension resizes and fills (correlation.py:22-27).  rbot1/rbot2 (the extension's padded
    NHWC re-layouts) are not needed by this implementation and are left empty."""

    @staticmethod
    def forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        _check(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        b, _, h, w = input1.shape
        output.resize_(b, 81, h, w)
        ops.corr81_fwd(input1, input2, out=output)
        return 1

    @staticmethod
    def backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size, kernel_size,
                 max_displacement, stride1, stride2, corr_multiply):
        _check(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        g1, g2 = ops.corr81_bwd(input1, input2, grad_output)
        grad_input1.resize_(g1.shape).copy_(g1)
        grad_input2.resize_(g2.shape).copy_(g2)
        return 1


class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        _check(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        ctx.save_for_backward(input1, input2)
        with torch.cuda.device_of(input1):
            return ops.corr81_fwd(input1, input2)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        with torch.cuda.device_of(input1):
            g1, g2 = ops.corr81_bwd(input1, input2, grad_output.contiguous())
        return g1, g2, None, None, None, None, None, None
