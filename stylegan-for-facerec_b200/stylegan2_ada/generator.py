"""B200 host mirror of restyle-encoder/models/stylegan2_ada/generator.py: same class names, constructor and
forward signatures, module tree and state_dict keys (a reference checkpoint loads with strict=True); the
arithmetic runs on the sg2_b200 kernels.  Selected by psp.py:24-30 (`opts.generator_ada`).

Per layer (SynthesisLayer2.forward, generator.py:186-204) the reference runs: affine -> per-sample weights
[N,O,I,k,k] -> demod -> grouped conv -> SmoothUpsample (3 ops) -> + noise -> + bias -> lrelu -> gain -> clamp.
Here: one modulation kernel (style + demod table), one shared-weight conv over the whole batch, and ONE
epilogue kernel (SmoothUpsample fused with noise, bias, lrelu, gain, clamp; or the plain epilogue when the
layer does not up-sample).  When a gradient is needed the layers compose the differentiable pieces instead
(utils.py).  Only the 'stylegan2' synthesis layer is implemented (the reference's default and the one psp.py builds)."""
import os

import numpy as np
import torch

from ..stylegan2 import functional as K
from ..stylegan2.op import fused_leaky_relu
from .utils import (FullyConnectedLayer, SmoothUpsample, SmoothUpsampleFunction, ada_bias_act, identity, modulated_conv2d,
                    normalize_2nd_moment, smooth_upsample2x)


class Generator(torch.nn.Module):                            # generator.py:6-52

    def __init__(self, z_dim, w_dim, w_num_layers, img_resolution, img_channels, synthesis_layer='stylegan2'):
        super().__init__()
        self.z_dim = z_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          synthesis_layer=synthesis_layer)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, w_dim=w_dim, num_ws=self.num_ws, num_layers=w_num_layers)

    def forward(self, z, truncation_psi=1, truncation_cutoff=None, noise_mode='random', input_is_latent=False,
                randomize_noise=None, return_latents=False):
        noise_mode = 'random' if randomize_noise else 'const'         # generator.py:23-26
        z = z[0]
        if not input_is_latent:
            ws = self.mapping(z, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
            out_synth = self.synthesis(ws, noise_mode, return_latents=return_latents)
        else:
            out_synth = self.synthesis(z, noise_mode, return_latents=return_latents)
        if return_latents:
            return out_synth[0], z
        return out_synth[0], None

    @property
    def precision(self):
        return self.synthesis.precision

    @precision.setter
    def precision(self, value):
        self.synthesis.precision = value

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.w_dim, device=self.synthesis.first_block.const.device)
        return self.mapping(latent_in, truncation_psi=1, truncation_cutoff=None).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.mapping(input, truncation_psi=1, truncation_cutoff=None)


class SynthesisNetwork(torch.nn.Module):                     # generator.py:55-88
    """The decoder proper: a 4x4 prologue (learned constant -> conv1 -> ToRGB) and one block per octave (conv0 with the
    2x SmoothUpsample, conv1, ToRGB added to the up-sampled running image).  ws [B, num_ws, w_dim] is cut into windows of
    3 rows that overlap by one (the ToRGB row of an octave is also the conv0 row of the next), as in the reference.

    Per call this issues 3 kernels per layer without autograd (modulation + demodulation table, shared-weight conv, fused
    epilogue) and the differentiable composition of utils.py / tc_route.py with it; `precision` selects the fp32 SIMT
    convolutions ('auto' / 'exact') or the tcgen05 kernel with bf16 operands ('bf16', or SG2_B200_PRECISION=bf16) --
    every 3x3 convolution of this decoder is stride 1, so all of them qualify."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=16384, channel_max=512, synthesis_layer='stylegan2'):
        super().__init__()
        log2 = int(np.log2(img_resolution))
        octaves = [4 << i for i in range(log2 - 1)]                      # 4, 8, ..., img_resolution
        width = {r: min(channel_base // r, channel_max) for r in octaves}
        self.w_dim, self.img_channels = w_dim, img_channels
        self.img_resolution, self.img_resolution_log2 = img_resolution, log2
        self.block_resolutions = octaves
        self.num_ws = 2 * len(octaves) + 2
        self.precision = os.environ.get('SG2_B200_PRECISION', 'auto')
        common = dict(w_dim=w_dim, img_channels=img_channels, synthesis_layer=synthesis_layer)
        self.blocks = torch.nn.ModuleList()                                # registered before first_block, like the reference
        self.first_block = SynthesisPrologue(width[4], resolution=4, **common)
        self.blocks.extend(SynthesisBlock(width[r // 2], width[r], resolution=r, **common) for r in octaves[1:])

    @property
    def precision(self):
        return self.__dict__.get('_precision', 'auto')

    @precision.setter
    def precision(self, value):
        if value not in ('auto', 'exact', 'bf16'):
            raise ValueError(f"sg2_b200 SynthesisNetwork.precision must be 'auto', 'exact' or 'bf16', got {value!r}")
        self.__dict__['_precision'] = value

    def invalidate_caches(self):
        """drop the frozen-weight packs of the tensor-core route (needed after an in-place `.data` write to a weight)"""
        for m in self.modules():
            m.__dict__.pop('_tc_cache', None)

    # -- whole-network bf16 engine (csrc/synth.cu, sg2_synth_create_ada): the same launch plan as the rosinality decoder's --
    def _use_engine(self, ws, noise_mode):
        if self.precision != 'bf16' or not ws.is_cuda or noise_mode not in ('const', 'random'):
            return False
        if K.needs_grad(ws, *self.parameters()) or os.environ.get('SG2_B200_ADA_ENGINE', '1') == '0':
            return False
        widths = [self.first_block.conv1.weight.shape[0]] + [b.conv1.weight.shape[0] for b in self.blocks]
        return (self.img_channels == 3 and self.w_dim % 32 == 0 and self.w_dim <= 512 and self.img_resolution >= 8
                and all(c % 16 == 0 and 16 <= c <= 512 for c in widths) and self.first_block.const.dtype == torch.float32)

    def engine(self):
        dev = self.first_block.const.device
        eng = self.__dict__.get('_engine')
        if eng is None or eng.device != dev:
            from ..engine import AdaSynthesisEngine
            eng = AdaSynthesisEngine(self)
            self.__dict__['_engine'] = eng
        return eng

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_engine', None)
        return state

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__.pop('_engine', None)
        return replica

    def forward(self, ws, noise_mode='random', return_latents=False, **kwargs):
        if self._use_engine(ws, noise_mode):
            eng = self.engine()
            noise = [None] * eng.G.num_layers
            if noise_mode == 'const':
                noise = [self.first_block.conv1.noise_const]
                for b in self.blocks:
                    noise += [b.conv0.noise_const, b.conv1.noise_const]
            img = eng.synthesize(ws[:, :eng.G.n_latent], noise)
            return (img, None) if return_latents else [img]
        split_ws = [ws[:, 0:2, :]] + [ws[:, 2 * n + 1: 2 * n + 4, :] for n in range(len(self.block_resolutions))]
        with K.tc_grad(True if self.precision == 'bf16' else (False if self.precision == 'exact' else None)):
            x, img = self.first_block(split_ws[0], noise_mode)
            for i in range(len(self.block_resolutions) - 1):
                x, img = self.blocks[i](x, img, split_ws[i + 1], noise_mode)
        if return_latents:
            return img, None
        return [img]


def _layer_classes(synthesis_layer):
    if synthesis_layer != 'stylegan2':
        raise NotImplementedError("sg2_b200 stylegan2_ada: only synthesis_layer='stylegan2' is implemented "
                                  "(the reference's default and the decoder psp.py builds)")
    return SynthesisLayer2, ToRGBLayer2


class SynthesisPrologue(torch.nn.Module):                    # generator.py:91-111
    """learned constant [C, 4, 4] -> conv1 (rows 0 of ws) -> ToRGB (row 1): the first feature map and the first image"""

    def __init__(self, out_channels, w_dim, resolution, img_channels, synthesis_layer):
        super().__init__()
        SynthesisLayer, ToRGBLayer = _layer_classes(synthesis_layer)
        self.w_dim, self.resolution, self.img_channels = w_dim, resolution, img_channels
        self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution)
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim)

    def forward(self, ws, noise_mode):
        w_iter = iter(ws.unbind(dim=1))
        x = self.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        x = self.conv1(x, next(w_iter), noise_mode=noise_mode)
        img = self.torgb(x, next(w_iter))
        return x, img


class SynthesisBlock(torch.nn.Module):                       # generator.py:114-139
    """one octave: conv0 (3x3 at the input resolution, then SmoothUpsample x2 fused with its epilogue), conv1, ToRGB; the
    running image is up-sampled by the same filter and the new ToRGB output added in the same pass"""

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, synthesis_layer):
        super().__init__()
        SynthesisLayer, ToRGBLayer = _layer_classes(synthesis_layer)
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.num_conv = self.num_torgb = 0                                  # attributes of the reference class, unused here too
        self.resampler = SmoothUpsample()
        self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, resampler=self.resampler)
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution)
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim)

    def forward(self, x, img, ws, noise_mode):
        w_iter = iter(ws.unbind(dim=1))
        x = self.conv0(x, next(w_iter), noise_mode=noise_mode)
        x = self.conv1(x, next(w_iter), noise_mode=noise_mode)
        y = self.torgb(x, next(w_iter))
        # img = resampler(img); img.add_(y)  (generator.py:135-137) in one pass
        if K.needs_grad(img, y):
            img = SmoothUpsampleFunction.apply(img, self.resampler.kernel) + y
        else:
            img = smooth_upsample2x(img, self.resampler.kernel, addend=y)
        return x, img


class ToRGBLayer2(torch.nn.Module):                          # generator.py:142-155

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1):
        super().__init__()
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def forward(self, x, w):
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        if K.needs_grad(x, w, self.weight, self.bias, self.affine.weight, self.affine.bias):
            styles = self.affine(w) * self.weight_gain
            y = modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False)
            return torch.clamp(y + self.bias.to(y.dtype)[None, :, None, None], -256, 256)
        # styles * weight_gain folded into the weights; no demodulation
        wt, _ = K.conv_prep(self.weight.to(x.dtype), self.weight_gain, want_wsq=False)
        s, _ = K.modulation(w.to(x.dtype), self.affine.weight, self.affine.bias, None, self.weight.shape[0],
                            self.affine.weight_gain, self.affine.bias_gain, False)
        y = K.shared_conv(x, wt, s, None, self.weight.shape[0], self.weight.shape[-1], 0)
        return ada_bias_act(y, None, None, self.bias, act=1, gain=1.0, clamp=256.0)


class SynthesisLayer2(torch.nn.Module):                      # generator.py:172-204

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, resampler=identity, activation='lrelu'):
        super().__init__()
        if activation != 'lrelu' or kernel_size != 3:
            raise NotImplementedError("sg2_b200 SynthesisLayer2: 3x3 / lrelu only")
        self.resolution = resolution
        self.resampler = resampler
        self.activation_gain = float(np.sqrt(2))
        self.padding = kernel_size // 2
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.register_buffer('noise_const', torch.randn([resolution, resolution]))
        self.noise_strength = torch.nn.Parameter(torch.zeros([1]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode, gain=1):
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        cout = self.weight.shape[0]
        noise = None
        if noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device, dtype=x.dtype)
        if noise_mode == 'const':
            noise = self.noise_const
        g, c = self.activation_gain * gain, 256.0 * gain
        if K.needs_grad(x, w, self.weight, self.bias, self.noise_strength, self.affine.weight, self.affine.bias):
            if not isinstance(self.resampler, SmoothUpsample) and K.tc_conv_ok(x, self.weight, 0):
                # the whole layer in three passes each way on the tensor-core route (functional.StyledConvTCFunction)
                s = self.affine(w).float()
                wt = self.weight
                if wt.requires_grad:
                    w4, wp, wp_adj, wsq = wt, None, None, wt.float().square().sum([2, 3])
                else:                             # frozen decoder: packed weights cached until the parameter changes
                    w4, wp, wp_adj, wsq = K.cached_tc_packs(self, lambda: wt, (wt.data_ptr(), wt._version, str(wt.device)), 0)
                d = torch.rsqrt(s.square() @ wsq.t() + 1e-8)
                nz = None if noise is None else noise.reshape(-1, 1, self.resolution, self.resolution)
                y = K.StyledConvTCFunction.apply(x, s, d, w4, nz, self.noise_strength, self.bias, 0.2, g, wp, wp_adj)
                return torch.clamp(y, -c, c)
            y = modulated_conv2d(x=x, weight=self.weight, styles=self.affine(w), padding=self.padding)
            if isinstance(self.resampler, SmoothUpsample):
                y = SmoothUpsampleFunction.apply(y, self.resampler.kernel)
            if noise is not None:
                y = y + noise.to(y.dtype) * self.noise_strength.to(y.dtype)
            return torch.clamp(fused_leaky_relu(y, self.bias.to(y.dtype), 0.2, g), -c, c)
        wt, wsq = K.conv_prep(self.weight.to(x.dtype), 1.0, want_wsq=True)
        s, d = K.modulation(w.to(x.dtype), self.affine.weight, self.affine.bias, wsq, cout, self.affine.weight_gain,
                            self.affine.bias_gain, True)
        if K.tc_conv_ok(x, self.weight, 0):       # tensor-core conv: modulation on the way in, demodulation in its epilogue
            y = K.tc_conv3x3(x * s.to(x.dtype).view(x.shape[0], -1, 1, 1), self.weight, scale=d)
        else:
            y = K.shared_conv(x, wt, s, d, cout, 3, 0)
        if isinstance(self.resampler, SmoothUpsample):
            return smooth_upsample2x(y, self.resampler.kernel, noise, self.noise_strength, self.bias, None, act=3,
                                     gain=g, clamp=c)
        return ada_bias_act(y, noise, self.noise_strength, self.bias, act=3, gain=g, clamp=c)


class MappingNetwork(torch.nn.Module):                       # generator.py:242-286

    def __init__(self, z_dim, w_dim, num_ws, num_layers=8, activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        self.z_dim = z_dim
        self.w_dim = w_dim
        self.num_ws = num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        self.lr_multiplier = lr_multiplier
        self.activation = activation
        features_list = [z_dim] + [w_dim] * num_layers
        self.layers = torch.nn.ModuleList()
        for idx in range(num_layers):
            self.layers.append(FullyConnectedLayer(features_list[idx], features_list[idx + 1], activation=activation,
                                                   lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        if not z.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")
        fused_ok = (getattr(self, 'activation', 'lrelu') == 'lrelu' and self.z_dim == self.w_dim and 0 < self.num_layers <= 32
                    and self.w_dim in (32, 64, 128, 256, 512) and z.dim() == 2)      # what sg2_mapping_fwd implements
        if fused_ok and not K.needs_grad(z, *self.parameters()):
            # normalize_2nd_moment + the whole MLP on the mapping kernels (one launch per layer)
            x = K.mapping(z, [l.weight for l in self.layers], [l.bias for l in self.layers], self.lr_multiplier, True)
        else:
            x = normalize_2nd_moment(z)
            for idx in range(self.num_layers):
                x = self.layers[idx](x)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x
