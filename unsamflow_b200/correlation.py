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


class CostConcatFunction(Function):
    """LeakyReLU(Correlation(a, b)) for one or two feature pairs, concatenated with further tensors along the channel
    axis -- models/pwclite.py:307-308 (+ :356-357 with add_mask_corr) and the torch.cat at :364-366 -- with the cost
    volumes written by the correlation kernel straight into their channel slice of the result, activation applied in
    the kernel's epilogue (usf_corr_fwd_into + USF_FLAG_LEAKY_RELU): no separate activation
    pass, no copy of the 81-channel volumes.  apply(slope, n_pairs, a_1, b_1, [a_2, b_2,] tail_1, ...)."""

    @staticmethod
    def forward(ctx, slope, n_pairs, *tensors):
        pairs, tail = tensors[:2 * n_pairs], tensors[2 * n_pairs:]
        x = pairs[0]
        B, _, H, W = x.shape
        K = ops.NUM_DISPLACEMENTS
        widths = [K] * n_pairs + [t.shape[1] for t in tail]
        buf = torch.empty((B, sum(widths), H, W), dtype=x.dtype, device=x.device)
        for i in range(n_pairs):
            ops.corr_fwd(pairs[2 * i], pairs[2 * i + 1], leaky_slope=slope, out=buf[:, i * K:(i + 1) * K])
        off = n_pairs * K
        for t in tail:
            buf[:, off:off + t.shape[1]].copy_(t)
            off += t.shape[1]
        ctx.slope, ctx.n_pairs, ctx.widths = slope, n_pairs, widths
        # the result is the next convolution's input, which autograd keeps anyway: saving it costs nothing extra
        ctx.save_for_backward(buf, *pairs)
        return buf

    @staticmethod
    @once_differentiable
    def backward(ctx, gbuf):
        buf, *pairs = ctx.saved_tensors
        K = ops.NUM_DISPLACEMENTS
        grads = []
        for i in range(ctx.n_pairs):
            sl = slice(i * K, (i + 1) * K)
            # LeakyReLU backward from the sign of its output (slope > 0 keeps the sign)
            g = torch.where(buf[:, sl] > 0, gbuf[:, sl], gbuf[:, sl] * ctx.slope)
            g1, g2 = ops.corr_bwd(pairs[2 * i], pairs[2 * i + 1], g, ctx.needs_input_grad[2 + 2 * i],
                                  ctx.needs_input_grad[3 + 2 * i])
            grads += [g1, g2]
        off = ctx.n_pairs * K
        for j, wdt in enumerate(ctx.widths[ctx.n_pairs:]):
            grads.append(gbuf[:, off:off + wdt] if ctx.needs_input_grad[2 + 2 * ctx.n_pairs + j] else None)
            off += wdt
        return (None, None, *grads)


def cost_concat(pairs, tail, slope=0.1):
    """cat([LeakyReLU(corr(a, b)) for (a, b) in pairs] + tail, dim=1) with the correlation kernels writing into the
    result (see CostConcatFunction)."""
    flat = [t for ab in pairs for t in ab]
    return CostConcatFunction.apply(float(slope), len(pairs), *flat, *tail)


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
