"""stylegan-for-facerec_b200 -- B200-native (sm_100a) StyleGAN2 synthesis path, a drop-in behind the
reference's `Generator` / `upfirdn2d` / `fused_leaky_relu` API.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it with

    import importlib; sg2 = importlib.import_module("stylegan-for-facerec_b200")

or through the `sg2_b200` shim module at the repo root (`import sg2_b200 as sg2`).
Layout mirrors the reference: `<pkg>.stylegan2.model`, `<pkg>.stylegan2.op`.
"""
from . import _lib
from .stylegan2 import model, op
from .stylegan2.model import (Blur, ConstantInput, ConvLayer, Discriminator, Downsample, EqualConv2d,
                              EqualLinear, Generator, ModulatedConv2d, NoiseInjection, PixelNorm, ResBlock,
                              ScaledLeakyReLU, StyledConv, ToRGB, Upsample, make_kernel)
from .stylegan2.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from . import psp_io, stylegan2_ada                      # ADA decoder variant, face_pool / resize / uint8 after the decoder

__all__ = ["Generator", "Discriminator", "ModulatedConv2d", "StyledConv", "ToRGB", "Blur", "Upsample",
           "Downsample", "EqualLinear", "EqualConv2d", "NoiseInjection", "ConstantInput", "PixelNorm",
           "ScaledLeakyReLU", "ConvLayer", "ResBlock", "FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d",
           "make_kernel", "model", "op", "psp_io", "stylegan2_ada", "_lib"]
__version__ = "0.1.0"
