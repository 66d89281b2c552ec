"""Drop-in `Correlation` module (same constructor and call signature as the reference).

Replaces models/correlation_package/correlation.py:75-102 (module) and :7-72
(CorrelationFunction), and is interchangeable with models/correlation_native.py:6-23.
models/pwclite.py:208-215 constructs it with keyword arguments
(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1) and calls
self.corr(x1, x2_warp) -> [B,81,H,W] (:307, :356).
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class CorrelationFunction(Function):
    """Saves the INPUTS only (like correlation.py:11): the caller mutates the output in place
    (LeakyReLU(inplace=True), models/pwclite.py:308)."""

    @staticmethod
    def forward(ctx, input1, input2):
        ctx.save_for_backward(input1, input2)
        return ops.corr_fwd(input1, input2)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.corr_bwd(input1, input2, grad_output, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g1, g2


class Correlation(nn.Module):
    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                 corr_multiply=1, *args, **kwargs):
        super().__init__()
        # The reference CUDA package takes general values but PWC-Lite only ever uses this one
        # configuration; anything else is rejected loudly instead of computing something else.
        if (pad_size, kernel_size, max_displacement, stride1, stride2) != (4, 1, 4, 1, 1):
            raise NotImplementedError(
                "unsamflow_b200.Correlation implements the PWC-Lite configuration only: pad_size=4, "
                f"kernel_size=1, max_displacement=4, stride1=1, stride2=1 (got {pad_size}, {kernel_size}, "
                f"{max_displacement}, {stride1}, {stride2})")
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply  # accepted and unused, as in correlation_cuda.cc
        self.output_dim = 2 * max_displacement + 1

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2)

    def extra_repr(self):
        return f"max_displacement={self.max_displacement}, out_channels={self.output_dim ** 2}"
