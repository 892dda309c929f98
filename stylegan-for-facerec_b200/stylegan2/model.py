"""StyleGAN2 config-f generator (and the discriminator-side consumers of the op API) on the sm_100a
kernels of libsg2_b200 -- a drop-in for the reference's model.py.

Same class names, constructor/forward signatures, attributes and ``state_dict`` keys/shapes as
/root/reference/backbone/stylegan2/model.py (== restyle-encoder/models/stylegan2/model.py), so
rosinality ``g_ema`` checkpoints load unchanged (psp.py:68-71, e4e.py:56-59).  What runs
underneath is different:

* inference (no autograd) in fp32/fp16 storage -> the exact NCHW kernels: one SIMT implicit-GEMM
  modulated conv with the weight operand shared by the batch (no [B,Cout,Cin,k,k] weights, no
  groups=B), smem FIR stencils, fused noise+bias+lrelu, fused ToRGB tail, one-launch mapping MLP;
* inference with ``precision='bf16'`` (or bfloat16 parameters) -> the whole-network engine
  (``engine.SynthesisEngine``): NHWC bf16 activations, tcgen05/TMEM implicit GEMM fed by TMA with
  demod/noise/bias/lrelu/next-layer-modulation/ToRGB fused into the epilogue;
* training (autograd on) -> the same kernels behind autograd Functions (dL/dx, dL/dstyle on the sg2
  conv kernel; weight gradients through the library wgrad).
"""
import math
import os
import random

import torch
from torch import nn
from torch.nn import functional as F

from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from . import functional as K


class PixelNorm(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, input):
        # model.py:14-15; stand-alone use only -- inside Generator.style it is fused into the mapping kernel
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _up_pad(n_taps, factor):            # model.py:37-42
    p = n_taps - factor
    return (p + 1) // 2 + factor - 1, p // 2


def _down_pad(n_taps, factor):          # model.py:58-63
    p = n_taps - factor
    return (p + 1) // 2, p // 2


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel) * (factor ** 2))
        self.pad = _up_pad(self.kernel.shape[0], factor)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer('kernel', make_kernel(kernel))
        self.pad = _down_pad(self.kernel.shape[0], factor)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer('kernel', kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """Equalised-lr plain convolution (discriminator side; a plain library conv, as in the reference)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return F.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride,
                        padding=self.padding)

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},'
                f' {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})')


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if input.is_cuda and not K.needs_grad(input, self.weight, self.bias) and \
                (self.bias is not None or not self.activation):
            return K.equal_linear(input, self.weight, self.bias, self.scale, self.lr_mul,
                                  bool(self.activation))
        # autograd path: library GEMM + the sg2 fused bias/activation op (model.py:147-155)
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        bias = None if self.bias is None else self.bias * self.lr_mul
        return F.linear(input, self.weight * self.scale, bias=bias)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ScaledLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, negative_slope=self.negative_slope) * math.sqrt(2)


class MappingNetwork(nn.Sequential):
    """``Generator.style``: still an nn.Sequential(PixelNorm, EqualLinear x n_mlp) -- same keys
    ``style.{1..n}.{weight,bias}`` -- but evaluated as ONE kernel when autograd is off."""

    def forward(self, input):
        layers = list(self)
        fusable = (input.is_cuda and input.dim() == 2 and len(layers) >= 1 and isinstance(layers[0], PixelNorm)
                   and all(isinstance(m, EqualLinear) and m.activation and m.bias is not None
                           and m.weight.shape == (input.shape[1], input.shape[1])
                           and m.lr_mul == layers[1].lr_mul for m in layers[1:])
                   and input.shape[1] in (32, 64, 128, 256, 512) and len(layers) <= 33)
        if fusable and not K.needs_grad(input, *self.parameters()):
            lr_mul = layers[1].lr_mul if len(layers) > 1 else 1.0
            return K.mapping(input, [m.weight for m in layers[1:]], [m.bias for m in layers[1:]], lr_mul, True)
        return super().forward(input)


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:        # model.py:198-204
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:      # model.py:206-212
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    @property
    def _mode(self):
        return 1 if self.upsample else (2 if self.downsample else 0)

    def _forward_kernels(self, input, style):
        """no-autograd path: prep -> modulation affine + demod table -> one shared-weight conv."""
        if self.kernel_size not in (1, 3):
            raise RuntimeError(f"sg2_b200 ModulatedConv2d: kernel_size {self.kernel_size} not in (1, 3)")
        w4 = self.weight[0]
        mod = self.modulation
        if self.kernel_size == 1 and self._mode == 0 and not self.demodulate and self.out_channel <= 4 and self.in_channel <= 1024:
            # ToRGB: one HBM-bound pass over the activation (the implicit-GEMM tile would be 3 rows of 64 tall)
            s, _ = K.modulation(style.to(input.dtype), mod.weight, mod.bias, None, self.out_channel, mod.scale, mod.lr_mul, False)
            return K.rgb_modconv(input, s, w4.reshape(self.out_channel, self.in_channel).float() * self.scale)
        wt, wsq = K.conv_prep(w4.to(input.dtype), self.scale, want_wsq=self.demodulate)
        s, d = K.modulation(style.to(input.dtype), mod.weight, mod.bias, wsq, self.out_channel, mod.scale,
                            mod.lr_mul, self.demodulate)
        if self.downsample:
            input = self.blur(input)
        out = K.shared_conv(input, wt, s, d, self.out_channel, self.kernel_size, self._mode)
        if self.upsample:
            out = self.blur(out)
        return out

    def _tc_weights(self, w4=None):
        """(w4, packed forward weights, packed adjoint weights, wsq [Cout,Cin]) for the tensor-core route; cached while the
        weight is frozen (the ReStyle direction), recomputed -- and differentiable through w4 / wsq -- when it trains"""
        w = self.weight
        if not (w.requires_grad and torch.is_grad_enabled()):
            return K.cached_tc_packs(self, lambda: w[0] * self.scale, (w.data_ptr(), w._version, str(w.device)), self._mode)
        if w4 is None:
            w4 = w[0] * self.scale
        return w4, None, None, w4.float().pow(2).sum([2, 3])

    def _forward_autograd(self, input, style):
        """differentiable path.  Same algebra as the kernels: y = d[b,co] * conv(W*scale, s[b,ci]*x),
        d = rsqrt(sum_ci s^2 * sum_k (scale*W)^2 + eps)  ==  model.py:236-240 without per-sample weights."""
        batch = input.shape[0]
        s = self.modulation(style)                                        # [B, Cin]
        w4 = self.weight[0] * self.scale                                   # [Cout, Cin, k, k]
        if self.kernel_size == 1 and self._mode == 0 and not self.demodulate and self.out_channel <= 4 and self.in_channel <= 1024:
            # ToRGB: one pass over the activation each way (functional.RgbModConvFunction), exact fp32 math
            return K.RgbModConvFunction.apply(input, s, w4.reshape(self.out_channel, self.in_channel))
        if K.tc_conv_ok(input, w4, self._mode):
            # tensor-core route: modulate + layout / conv / demodulate + layout, each one pass, grad_s and grad_d
            # produced by the adjoint passes (functional.ModulatedConvTCFunction)
            w4, wp, wp_adj, wsq = self._tc_weights(w4)
            d = torch.rsqrt(s.float().pow(2) @ wsq.t() + self.eps) if self.demodulate else None
            y = K.ModulatedConvTCFunction.apply(input, s.float(), d, w4, self._mode, wp, wp_adj)
            return self.blur(y) if self.upsample else y
        x = input * s.to(input.dtype).view(batch, -1, 1, 1)
        if self.downsample:
            x = self.blur(x)
        y = K.SharedConvFunction.apply(x, w4.to(input.dtype), self._mode)
        if self.demodulate:
            wsq = w4.float().pow(2).sum([2, 3])                            # [Cout, Cin]
            d = torch.rsqrt(s.float().pow(2) @ wsq.t() + self.eps)          # [B, Cout]
            y = y * d.to(y.dtype).view(batch, -1, 1, 1)
        if self.upsample:
            y = self.blur(y)
        return y

    def forward(self, input, style):
        if not input.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        if K.needs_grad(input, style, self.weight, self.modulation.weight, self.modulation.bias):
            return self._forward_autograd(input, style)
        return self._forward_kernels(input, style)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        if image.is_cuda and not K.needs_grad(image, noise, self.weight):
            return K.noise_bias_act(image, noise, self.weight, None, act=1, act_scale=1.0)
        return image + self.weight * noise


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False,
                 blur_kernel=[1, 3, 3, 1], demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        conv = self.conv
        if (input.is_cuda and conv._mode == 0 and conv.kernel_size == 3 and K.tc_conv_ok(input, conv.weight[0], 0)
                and K.needs_grad(input, style, noise, *self.parameters())):
            # differentiable tensor-core route: the whole layer in three passes each way (functional.StyledConvTCFunction)
            if noise is None:
                noise = input.new_empty(input.shape[0], 1, input.shape[2], input.shape[3]).normal_()
            s = conv.modulation(style)
            w4, wp, wp_adj, wsq = conv._tc_weights()
            d = torch.rsqrt(s.float().pow(2) @ wsq.t() + conv.eps) if conv.demodulate else None
            return K.StyledConvTCFunction.apply(input, s.float(), d, w4, noise, self.noise.weight, self.activate.bias,
                                                self.activate.negative_slope, self.activate.scale, wp, wp_adj)
        out = self.conv(input, style)
        if out.is_cuda and K.tc_grad_enabled() and K.needs_grad(out, noise, self.noise.weight, self.activate.bias):
            # tensor-core route, after the blur of an up-sampling layer: noise + bias + activation in one pass each way
            if noise is None:
                noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
            return K.NoiseBiasActFunction.apply(out, noise, self.noise.weight, self.activate.bias,
                                                self.activate.negative_slope, self.activate.scale)
        if not K.needs_grad(out, noise, self.noise.weight, self.activate.bias):
            # conv -> (+ w*noise) -> (+ bias) -> lrelu*sqrt(2) in one pass over the activation
            if noise is None:
                noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
            return K.noise_bias_act(out, noise, self.noise.weight, self.activate.bias, act=3,
                                    alpha=self.activate.negative_slope, act_scale=self.activate.scale)
        out = self.noise(out, noise=noise)
        return self.activate(out)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style)
        if not K.needs_grad(out, skip, self.bias):
            if skip is None:
                return K.torgb_combine(out, self.bias, None, None, (0, 0))
            up = self.upsample
            if up.factor == 2 and skip.shape[2] * 2 == out.shape[2] and skip.shape[3] * 2 == out.shape[3]:
                return K.torgb_combine(out, self.bias, skip, up.kernel, up.pad)
        out = out + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.style = MappingNetwork(*layers)
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                         128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                         512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        self._blur_kernel = list(blur_kernel)
        # 'auto': bf16 engine iff the parameters are bfloat16; 'exact': always the fp32-accumulate
        # NCHW kernels; 'bf16': tensor-core engine with fp32 master weights.  Env override for
        # unmodified callers: SG2_B200_PRECISION=bf16
        self.precision = os.environ.get('SG2_B200_PRECISION', 'auto')
        self._engine = None

    # -- engine / cache lifetime ---------------------------------------------------------------
    # The bf16 engine owns a ctypes plan handle, a device workspace and CUDA graphs: none of it may travel with a copy
    # of the module.  deepcopy / pickle / torch.save(G) drop it (the copy re-plans lazily on its first bf16 forward), and
    # nn.DataParallel replicas (whose parameters are fresh broadcast copies on every forward) get a transient engine
    # without CUDA graphs on their own device instead of sharing the device-0 one.
    @property
    def precision(self):
        return self.__dict__.get('_precision', 'auto')

    @precision.setter
    def precision(self, value):
        if value not in ('auto', 'exact', 'bf16'):
            raise ValueError(f"sg2_b200 Generator.precision must be 'auto', 'exact' or 'bf16', got {value!r}")
        self.__dict__['_precision'] = value

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_engine'] = None
        state.pop('_train_engine', None)
        return state

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__['_engine'] = None
        replica.__dict__.pop('_train_engine', None)
        replica.__dict__['_transient_engine'] = True
        return replica

    def invalidate_caches(self):
        """Drop every packed copy of the parameters (the engine's bf16 weight pack and CUDA graphs, the frozen-weight
        packs of the tensor-core autograd route).  Parameter updates through optimizers, `copy_`, `load_state_dict` or
        `.to()` are detected automatically (tensor version counter / storage address); an in-place write through `.data`
        bumps neither, so code that updates weights that way (rosinality-style EMA `accumulate()`) must call this."""
        self._engine = None
        self.__dict__.pop('_train_engine', None)
        for m in self.modules():
            m.__dict__.pop('_tc_cache', None)

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    # ------------------------------------------------------------------------------------------
    def _use_engine(self, latent, noise, return_features):
        mode = self.precision
        if mode == 'exact' or return_features or not latent.is_cuda:
            return False
        if K.needs_grad(latent, *self.parameters()) or any(n is not None and n.requires_grad for n in noise):
            return False
        if mode == 'bf16':
            return True
        return mode == 'auto' and self.input.input.dtype == torch.bfloat16

    def _use_train_engine(self, latent, noise, return_features):
        """the frozen-decoder fine-tuning direction (coach_restyle_psp.py:138-168): gradient w.r.t. the latents only"""
        if self.precision != 'bf16' or return_features or not latent.is_cuda or not torch.is_grad_enabled():
            return False
        if os.environ.get('SG2_B200_TRAIN_ENGINE', '1') == '0' or not latent.requires_grad:
            return False
        if any(p.requires_grad for p in self.parameters()) or any(n is not None and n.requires_grad for n in noise):
            return False
        if self.input.input.dtype != torch.float32 or self.style_dim % 32 or self.size < 8:
            return False
        chans = [self.channels[2 ** i] for i in range(2, self.log_size + 1)]
        return all(c >= 32 and c <= 512 and (c & (c - 1)) == 0 for c in chans) and list(self._blur_kernel) == [1, 3, 3, 1]

    def train_engine(self):
        dev = self.input.input.device
        te = self.__dict__.get('_train_engine')
        if te is None or te.device != dev:
            from ..engine import SynthesisEngine
            te = SynthesisEngine(self, training=True, use_graph=False)
            self.__dict__['_train_engine'] = te
        return te

    def engine(self):
        """The whole-network bf16 tcgen05 engine bound to this module's parameters (lazy, one per device: moving the
        module re-plans)."""
        dev = self.input.input.device
        if self._engine is None or self._engine.device != dev:
            from ..engine import SynthesisEngine
            self._engine = SynthesisEngine(self, use_graph=None if not self.__dict__.get('_transient_engine') else False)
        return self._engine

    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]
        # one z, no truncation, bf16 engine: the mapping network and the broadcast over the layers go into the engine's CUDA
        # graph with the synthesis (the arithmetic is the same: Generator.style on the same kernels)
        if (not input_is_latent and len(styles) == 1 and truncation >= 1 and torch.is_tensor(styles[0]) and styles[0].ndim == 2
                and styles[0].shape[1] == self.style_dim and self._use_engine(styles[0], noise, return_features)
                and not K.needs_grad(styles[0])):
            image, latent = self.engine().synthesize(None, noise, z=styles[0], want_latent=return_latents)
            return (image, latent) if return_latents else (image, None)
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            if styles[0].ndim < 3:
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            else:
                latent = styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
            latent = torch.cat([latent, latent2], 1)

        if self._use_engine(latent, noise, return_features):
            image = self.engine().synthesize(latent, noise)
            return (image, latent) if return_latents else (image, None)
        if self._use_train_engine(latent, noise, return_features):
            from ..engine import SynthesisFunction
            image = SynthesisFunction.apply(latent, self.train_engine(), noise)
            return (image, latent) if return_latents else (image, None)

        # precision == 'bf16' with gradients: the 3x3 stride-1 convolutions of the differentiable path (forward and
        # input gradient) run on the tensor-core kernel with bf16 operands; everything else stays fp32
        with K.tc_grad(True if self.precision == 'bf16' else (False if self.precision == 'exact' else None)):
            out = self.input(latent)
            out = self.conv1(out, latent[:, 0], noise=noise[0])
            skip = self.to_rgb1(out, latent[:, 1])
            i = 1
            for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2],
                                                            noise[2::2], self.to_rgbs):
                out = conv1(out, latent[:, i], noise=noise1)
                out = conv2(out, latent[:, i + 1], noise=noise2)
                skip = to_rgb(out, latent[:, i + 2], skip)
                i += 2
        image = skip
        if return_latents:
            return image, latent
        elif return_features:
            return image, out
        else:
            return image, None


# ------------------------------------------------------------------------------------------------
# discriminator-side consumers of the op API (model.py:545-673 of the reference): discriminator.py,
# instantiated over the building blocks above and re-exported here under the reference's names.
# ------------------------------------------------------------------------------------------------
from .discriminator import _make_classes, feature_channels, minibatch_stddev  # noqa: E402,F401

ConvLayer, ResBlock, Discriminator = _make_classes(Blur, EqualConv2d, EqualLinear, FusedLeakyReLU, ScaledLeakyReLU)
for _cls in (ConvLayer, ResBlock, Discriminator):      # picklable / printable under the module users import them from
    _cls.__module__ = __name__
    _cls.__qualname__ = _cls.__name__
