"""B200 host mirror of restyle-encoder/models/stylegan2_ada/utils.py (same names, arguments, state_dict
keys); the arithmetic runs on the sg2_b200 kernels (csrc/ada_ops.cu, modconv_simt.cu, linear.cu).

Two paths, like stylegan2/model.py: without autograd every layer is three fused kernels; when a gradient is needed
(the ReStyle coaches back-propagate through the frozen decoder) the same arithmetic is composed from the
differentiable pieces -- the shared-weight conv `Function` (forward and dgrad on the sg2 kernel), the
SmoothUpsample `Function` below (forward + its exact adjoint kernel), `fused_leaky_relu` -- with torch elementwise
ops in between."""
import math

import numpy as np
import torch

from .. import _lib
from ..stylegan2 import functional as K


def _no_grad_path(*tensors):
    if K.needs_grad(*tensors):
        raise RuntimeError("sg2_b200 stylegan2_ada: the ADA decoder kernels are inference-only (no autograd yet); "
                           "wrap the call in torch.no_grad() or freeze the decoder and detach its inputs")


def clamp_gain(x: torch.Tensor, g: float, c: float):       # utils.py:6-7
    return ada_bias_act(x, None, None, None, act=1, gain=g, clamp=c)


def normalize_2nd_moment(x, dim=1, eps=1e-8):                # utils.py:10-11 (== PixelNorm)
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def identity(x):
    return x


def leaky_relu_0_2(x):
    return torch.nn.functional.leaky_relu(x, 0.2)


activation_funcs = {"linear": {"fn": identity, "def_gain": 1}, "lrelu": {"fn": leaky_relu_0_2, "def_gain": np.sqrt(2)}}


def ada_bias_act(x, noise, noise_strength, bias, act=1, alpha=0.2, gain=1.0, clamp=0.0):
    """clamp(act(x + noise * noise_strength + bias[c]) * gain, -clamp, clamp) in one pass (generator.py:148-151,198-204)."""
    _lib.require_cuda(x)
    _no_grad_path(x, noise, noise_strength, bias)
    lib = _lib.load()
    x = x.contiguous()
    B, Cn = x.shape[0], x.shape[1]
    HW = int(math.prod(x.shape[2:]))
    nstride = 0
    if noise is not None:
        noise = noise.detach().to(x.dtype).contiguous()
        if noise.numel() == B * HW:
            nstride = HW
        elif noise.numel() != HW:
            raise RuntimeError(f"noise of shape {tuple(noise.shape)} does not broadcast to {tuple(x.shape)}")
        noise_strength = noise_strength.detach().to(x.dtype).contiguous()
    if bias is not None:
        bias = bias.detach().to(x.dtype).contiguous()
    out = torch.empty_like(x)
    with _lib.device_of(x):
        _lib.check(lib.sg2_ada_bias_act(out.data_ptr(), x.data_ptr(), _lib.ptr(noise), nstride,
                                        _lib.ptr(noise_strength) if noise is not None else None, _lib.ptr(bias), B, Cn, HW,
                                        act, float(alpha), float(gain), float(clamp), _lib.dtype_code(x),
                                        _lib.stream_of(x)), "ada_bias_act")
    return out


def smooth_upsample2x(x, kernel, noise=None, noise_strength=None, bias=None, addend=None, act=1, alpha=0.2, gain=1.0,
                      clamp=0.0):
    """SmoothUpsample (utils.py:76-95) fused with the epilogue that follows it (csrc/ada_ops.cu)."""
    _lib.require_cuda(x)
    _no_grad_path(x, noise, noise_strength, bias, addend)
    lib = _lib.load()
    x = x.contiguous()
    B, Cn, H, W = x.shape
    taps = kernel.detach().to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
    if taps.numel() != 16:
        raise RuntimeError("sg2_b200 SmoothUpsample: the kernel must have 4x4 taps")
    nstride = 0
    if noise is not None:
        noise = noise.detach().to(x.dtype).contiguous()
        if noise.numel() == B * 4 * H * W:
            nstride = 4 * H * W
        elif noise.numel() != 4 * H * W:
            raise RuntimeError(f"noise of shape {tuple(noise.shape)} does not broadcast to [{B}, 1, {2 * H}, {2 * W}]")
        noise_strength = noise_strength.detach().to(x.dtype).contiguous()
    if bias is not None:
        bias = bias.detach().to(x.dtype).contiguous()
    if addend is not None:
        addend = addend.detach().to(x.dtype).contiguous()
        if tuple(addend.shape) != (B, Cn, 2 * H, 2 * W):
            raise RuntimeError("sg2_b200 SmoothUpsample: addend shape mismatch")
    out = torch.empty((B, Cn, 2 * H, 2 * W), device=x.device, dtype=x.dtype)
    with _lib.device_of(x):
        _lib.check(lib.sg2_smooth_upsample2x(out.data_ptr(), x.data_ptr(), taps.data_ptr(), B, Cn, H, W, _lib.ptr(noise),
                                             nstride, _lib.ptr(noise_strength) if noise is not None else None,
                                             _lib.ptr(bias), _lib.ptr(addend), act, float(alpha), float(gain), float(clamp),
                                             _lib.dtype_code(x), _lib.stream_of(x)), "smooth_upsample2x")
    return out


class SmoothUpsampleFunction(torch.autograd.Function):
    """plain SmoothUpsample with autograd: forward sg2_smooth_upsample2x, backward its adjoint kernel"""

    @staticmethod
    def forward(ctx, x, kernel):
        ctx.save_for_backward(kernel)
        with torch.no_grad():
            return smooth_upsample2x(x.detach(), kernel.detach())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        kernel, = ctx.saved_tensors
        gy = gy.contiguous()
        b, c, oh, ow = gy.shape
        taps = kernel.detach().to(device=gy.device, dtype=torch.float32).reshape(-1).contiguous()
        gx = torch.empty((b, c, oh // 2, ow // 2), device=gy.device, dtype=gy.dtype)
        with _lib.device_of(gy):
            _lib.check(_lib.load().sg2_smooth_upsample2x_bwd(gx.data_ptr(), gy.data_ptr(), taps.data_ptr(), b * c, oh // 2,
                                                             ow // 2, _lib.dtype_code(gy), _lib.stream_of(gy)),
                       "smooth_upsample2x_bwd")
        return gx, None


class FullyConnectedLayer(torch.nn.Module):                  # utils.py:34-52
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.activation_name = activation
        self.activation = activation_funcs[activation]['fn']
        self.activation_gain = activation_funcs[activation]['def_gain']
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        lrelu = self.activation_name == 'lrelu'
        if not K.needs_grad(x, self.weight, self.bias) and (self.bias is not None or not lrelu):
            # act(x W^T * weight_gain + b * bias_gain) * act_gain == EqualLinear (linear.cu): lrelu carries the sqrt(2) gain
            return K.equal_linear(x, self.weight, self.bias, self.weight_gain, self.bias_gain, lrelu)
        # autograd path (mean_latent / get_latent are called with gradients enabled by the coaches): library GEMM +
        # the sg2 fused bias/activation op, like EqualLinear in stylegan2/model.py
        w = self.weight * self.weight_gain
        b = None if self.bias is None else self.bias * self.bias_gain
        if lrelu:
            from ..stylegan2.op import fused_leaky_relu
            out = torch.nn.functional.linear(x, w)
            return fused_leaky_relu(out, b if b is not None else torch.zeros(w.shape[0], device=x.device, dtype=x.dtype))
        return torch.nn.functional.linear(x, w, b)


class SmoothUpsample(torch.nn.Module):                       # utils.py:76-95
    def __init__(self):
        super().__init__()
        kernel = torch.tensor([[[[1, 3, 3, 1], [3, 9, 9, 3], [3, 9, 9, 3], [1, 3, 3, 1]]]], dtype=torch.float)
        kernel /= kernel.sum()
        self.kernel = torch.nn.Parameter(kernel, requires_grad=False)

    def forward(self, x: torch.Tensor):
        if K.needs_grad(x):
            return SmoothUpsampleFunction.apply(x, self.kernel)
        return smooth_upsample2x(x, self.kernel)


def modulated_conv2d(x, weight, styles, padding=0, demodulate=True):
    """utils.py:118-137 without per-sample weights: y = d[b,o] * conv(W, s[b,i] * x), d = rsqrt(sum (W s)^2 + 1e-8)."""
    _lib.require_cuda(x)
    k = weight.shape[-1]
    if k not in (1, 3) or padding != k // 2:
        raise RuntimeError(f"sg2_b200 modulated_conv2d: kernel {k} / padding {padding} not supported (1x1 pad 0, 3x3 pad 1)")
    if K.needs_grad(x, weight, styles):
        b = x.shape[0]
        if k == 1 and not demodulate and weight.shape[0] <= 4 and weight.shape[1] <= 1024:     # ToRGB: one pass each way
            return K.RgbModConvFunction.apply(x, styles, weight.reshape(weight.shape[0], weight.shape[1]))
        if K.tc_conv_ok(x, weight, 0):                                          # tensor-core route, see functional.py
            d = None
            if demodulate:
                d = torch.rsqrt(styles.float().square() @ weight.float().square().sum([2, 3]).t() + 1e-8)
            return K.ModulatedConvTCFunction.apply(x, styles.float(), d, weight, 0)
        y = K.SharedConvFunction.apply(x * styles.to(x.dtype).view(b, -1, 1, 1), weight.to(x.dtype), 0)
        if demodulate:
            wsq = weight.float().square().sum([2, 3])                          # [Cout, Cin]
            d = torch.rsqrt(styles.float().square() @ wsq.t() + 1e-8)           # [B, Cout]
            y = y * d.to(y.dtype).view(b, -1, 1, 1)
        return y
    wt, wsq = K.conv_prep(weight.to(x.dtype), 1.0, want_wsq=demodulate)
    s = styles.detach().float().contiguous()
    d = None
    if demodulate:
        d = torch.rsqrt(s.square() @ wsq + 1e-8).contiguous()      # [B, Cout]; tiny (B x Cin x Cout)
    return K.shared_conv(x, wt, s, d, weight.shape[0], k, 0)
