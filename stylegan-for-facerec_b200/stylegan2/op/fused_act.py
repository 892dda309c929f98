"""fused_leaky_relu / FusedLeakyReLU over the sm_100a kernel `sg2_fused_bias_act`.

Drop-in for the reference's op/fused_act.py (FusedLeakyReLUFunctionBackward :18-47,
FusedLeakyReLUFunction :50-69, FusedLeakyReLU :72-81, fused_leaky_relu :84-85): same names,
defaults, dtype preservation, first and second derivatives.  Extra over the reference: bfloat16,
64-bit element counts, no import-time JIT.
"""
import math

import torch
from torch import nn
from torch.autograd import Function

from ... import _lib


def bias_act(x, bias, ref, act, grad, alpha, scale):
    """out = act(x + bias[channel]) * scale through the C ABI (same 7 arguments as the reference's
    pybind `fused.fused_bias_act`, op/fused_bias_act.cpp:11-21).  bias / ref may be None or empty."""
    _lib.require_cuda(x)
    lib = _lib.load()
    x = x.contiguous()
    if bias is not None and bias.numel() == 0:
        bias = None
    if ref is not None and ref.numel() == 0:
        ref = None
    if bias is not None:
        _lib.require_cuda(bias, "bias")
        bias = bias.contiguous().to(x.dtype)
    if ref is not None:
        _lib.require_cuda(ref, "refer")
        ref = ref.contiguous().to(x.dtype)
        if ref.numel() != x.numel():
            raise RuntimeError("fused_bias_act: refer must have as many elements as input")
    out = torch.empty_like(x)
    step_b = int(math.prod(x.shape[2:])) if x.dim() > 2 else 1     # fused_bias_act_kernel.cu:67-71
    size_b = bias.numel() if bias is not None else 1
    with _lib.device_of(x):
        _lib.check(lib.sg2_fused_bias_act(out.data_ptr(), x.data_ptr(), _lib.ptr(bias), _lib.ptr(ref),
                                          x.numel(), step_b, size_b, act, grad, float(alpha), float(scale),
                                          _lib.dtype_code(x), _lib.stream_of(x)), "fused_bias_act")
    return out


def grad_bias_reduce(grad_input):
    """grad_bias[c] = sum of grad_input over every dim but 1 (op/fused_act.py:31-36)."""
    lib = _lib.load()
    g = grad_input.contiguous()
    B, Cn = g.shape[0], g.shape[1]
    HW = int(math.prod(g.shape[2:])) if g.dim() > 2 else 1
    out = torch.empty(Cn, device=g.device, dtype=torch.float32)
    with _lib.device_of(g):
        _lib.check(lib.sg2_bias_act_grad_bias(out.data_ptr(), g.data_ptr(), B, Cn, HW, _lib.dtype_code(g),
                                              _lib.stream_of(g)), "bias_act_grad_bias")
    return out.to(g.dtype)


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale, want_bias=True):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        # the reduction over the whole activation is skipped when the bias is frozen (the ReStyle direction)
        grad_bias = grad_bias_reduce(grad_input).detach() if want_bias else grad_input.new_zeros(out.shape[1])
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(
            grad_output, out, ctx.negative_slope, ctx.scale, ctx.needs_input_grad[1])
        return grad_input, (grad_bias if ctx.needs_input_grad[1] else None), None, None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
