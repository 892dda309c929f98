"""Discriminator-side consumers of the op API: `ConvLayer`, `ResBlock`, `Discriminator` with the
module tree, constructor arguments and state_dict keys of the reference's backbone/stylegan2/model.py:545-673, so that
third-party rosinality checkpoints and training code load unchanged.  The reference repo itself never instantiates them.

Everything here is composition: the blur of the down-sampling layers is the sg2 `upfirdn2d` kernel (through `Blur`), the
activation the sg2 `fused_leaky_relu` kernel (through `FusedLeakyReLU`), the convolutions the library conv the reference
uses too (`EqualConv2d`).  Re-exported by model.py.
"""
import math

import torch
from torch import nn


def feature_channels(resolution: int, channel_multiplier: int = 2) -> int:
    """channel count of the feature map at `resolution` (the table at model.py:549-559 / 390-400 as a formula):
    512 up to 32x32, then 16384 / resolution scaled by the multiplier."""
    return 512 if resolution <= 32 else (16384 // resolution) * channel_multiplier


def minibatch_stddev(x: torch.Tensor, group_size: int = 4, n_feat: int = 1) -> torch.Tensor:
    """The minibatch standard-deviation feature (model.py:655-668): the batch is cut into `batch / group` interleaved
    groups of `group` samples; per group and feature slice, the standard deviation over the group's samples averaged
    over channels and pixels becomes one extra constant feature map, appended to every sample of that group."""
    b, c, h, w = x.shape
    g = min(b, group_size)
    if b % g:
        raise RuntimeError(f"minibatch stddev: batch {b} is not a multiple of the group size {g}")
    grouped = x.reshape(g, b // g, n_feat, c // n_feat, h, w)
    centred = grouped - grouped.mean(dim=0, keepdim=True)
    std = (centred.square().mean(dim=0) + 1e-8).sqrt()                  # [b/g, n_feat, c/n_feat, h, w]
    feat = std.mean(dim=(2, 3, 4))                                      # [b/g, n_feat]
    feat = feat.reshape(1, b // g, n_feat, 1, 1).expand(g, -1, -1, h, w).reshape(b, n_feat, h, w)
    return torch.cat([x, feat.to(x.dtype)], dim=1)


def _make_classes(Blur, EqualConv2d, EqualLinear, FusedLeakyReLU, ScaledLeakyReLU):
    """the three classes, closed over the building blocks of model.py (which imports this module)"""

    class ConvLayer(nn.Sequential):
        """[Blur ->] EqualConv2d [-> FusedLeakyReLU | ScaledLeakyReLU]; children are positional, as in the reference, so
        keys read `0.kernel`, `1.weight`, `2.bias` with a blur and `0.weight`, `1.bias` without."""

        def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                     activate=True):
            stages = []
            if downsample:                       # blur, then a stride-2 VALID conv: together a 'same'-sized halving
                slack = len(blur_kernel) - 2 + kernel_size - 1
                stages.append(Blur(blur_kernel, pad=((slack + 1) // 2, slack // 2)))
            self.padding = 0 if downsample else kernel_size // 2
            stages.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding,
                                      stride=2 if downsample else 1, bias=bias and not activate))
            if activate:                         # the bias lives in the fused activation when there is one
                stages.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
            super().__init__(*stages)

    class ResBlock(nn.Module):
        def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
            super().__init__()
            self.conv1 = ConvLayer(in_channel, in_channel, 3)
            self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True)
            self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, activate=False, bias=False)

        def forward(self, input):
            residual = self.conv2(self.conv1(input))
            return (residual + self.skip(input)) * (1 / math.sqrt(2))

    class Discriminator(nn.Module):
        def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
            super().__init__()
            if size < 8 or size > 1024 or size & (size - 1):
                raise ValueError(f"Discriminator: size must be a power of two in [8, 1024], got {size}")
            resolutions = [size >> i for i in range(int(math.log2(size)) - 1)]            # size, size/2, ..., 4
            widths = [feature_channels(r, channel_multiplier) for r in resolutions]
            trunk = [ConvLayer(3, widths[0], 1)]
            trunk += [ResBlock(cin, cout, blur_kernel) for cin, cout in zip(widths, widths[1:])]
            self.convs = nn.Sequential(*trunk)
            self.stddev_group = 4
            self.stddev_feat = 1
            self.final_conv = ConvLayer(widths[-1] + 1, widths[-1], 3)
            self.final_linear = nn.Sequential(EqualLinear(widths[-1] * 4 * 4, widths[-1], activation='fused_lrelu'),
                                              EqualLinear(widths[-1], 1))

        def forward(self, input):
            feat = minibatch_stddev(self.convs(input), self.stddev_group, self.stddev_feat)
            return self.final_linear(self.final_conv(feat).flatten(1))

    return ConvLayer, ResBlock, Discriminator
