"""CPU oracle for the StyleGAN2 synthesis hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a *functional* CPU restatement (plain torch-CPU tensor arithmetic, no nn.Module, no
CUDA) of the algorithm the reference implements in

    /root/reference/backbone/stylegan2/model.py            (== restyle-encoder/models/stylegan2/model.py)
    /root/reference/backbone/stylegan2/op/fused_act.py     + op/fused_bias_act_kernel.cu
    /root/reference/backbone/stylegan2/op/upfirdn2d.py     + op/upfirdn2d_kernel.cu

Every function cites the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it -- and there only as
the checker / the reported CPU baseline.  The product package never imports this module and fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or known-answer fixtures for this path
("parity unpinned by the reference's own tests").  The pin used instead is
the reference itself, imported and executed in the authoring container through the CPU op shim of
BASELINE.md section 3 by ``tests/golden/make_golden.py``; its outputs are committed under
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` holds this restatement to them.

The dense contraction itself is third-party arithmetic absent from /root/reference:
``torch.nn.functional.conv2d / conv_transpose2d / linear`` (reference pin torch==1.6.0,
restyle-encoder/environment/restyle_env.yaml:34; here torch 2.11 CPU = oneDNN/MKL).  The oracle
calls the same library entry points at the same call sites as model.py:149,153,254,263,269.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

LRELU_SLOPE = 0.2
LRELU_GAIN = 2.0 ** 0.5

# Channel table of model.py:389-399 (channel_multiplier applies from 64x64 upward).
def channel_table(channel_multiplier: int = 2) -> Dict[int, int]:
    tab = {4: 512, 8: 512, 16: 512, 32: 512}
    base = {64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}
    for res, c in base.items():
        tab[res] = c * channel_multiplier
    return tab


# --------------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------------
def fir_kernel_2d(taps: Sequence[float]) -> Tensor:
    """model.py:18-26 make_kernel: outer product of a 1-D tap list, normalised to sum 1."""
    k = torch.as_tensor(list(taps), dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def bias_act(x: Tensor, bias: Optional[Tensor], ref: Optional[Tensor], act: int, grad: int,
             alpha: float, scale: float) -> Tensor:
    """fused_bias_act_kernel.cu:18-49.  bias indexes dim 1 (step_b = prod(shape[2:]), :69-71).

    act 1 = linear, act 3 = leaky-relu; grad 0 = value, 1 = first derivative keyed on ``ref``
    (the saved forward output), 2 = second derivative (identically zero)."""
    v = x
    if bias is not None and bias.numel() > 0:
        v = v + bias.reshape(1, -1, *([1] * (x.ndim - 2))).to(x.dtype)
    code = act * 10 + grad
    if code in (10, 11):
        y = v
    elif code in (12, 32):
        y = torch.zeros_like(v)
    elif code == 30:
        y = torch.where(v > 0, v, v * alpha)
    elif code == 31:
        y = torch.where(ref > 0, v, v * alpha)
    else:  # the CUDA switch falls into `default: case 10`
        y = v
    return y * scale


def fused_leaky_relu(x: Tensor, bias: Tensor, negative_slope: float = LRELU_SLOPE,
                     scale: float = LRELU_GAIN) -> Tensor:
    """op/fused_act.py:50-60,84-85 forward: act=3, grad=0."""
    return bias_act(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_backward(grad_out: Tensor, out: Tensor, negative_slope: float = LRELU_SLOPE,
                              scale: float = LRELU_GAIN) -> Tuple[Tensor, Tensor]:
    """op/fused_act.py:20-38: grad_input keyed on the saved output; grad_bias = sum over all dims but 1."""
    gi = bias_act(grad_out, None, out, 3, 1, negative_slope, scale)
    dims = [0] + list(range(2, gi.ndim))
    return gi, gi.sum(dims)


def upfirdn2d(x: Tensor, kernel: Tensor, up: int = 1, down: int = 1,
              pad: Tuple[int, int] = (0, 0)) -> Tensor:
    """op/upfirdn2d.py:142-147 (public call) over op/upfirdn2d.py:150-184 (the reference's own
    PyTorch statement of the op) == op/upfirdn2d_kernel.cu:52-137.

    x [B,C,H,W]; zero-insert upsample by ``up`` -> pad (negative pads crop) -> true 2-D convolution
    with ``kernel`` (i.e. correlation with the flipped taps) -> keep every ``down``-th sample."""
    return upfirdn2d_xy(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_xy(x: Tensor, kernel: Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                 px0: int, px1: int, py0: int, py1: int) -> Tensor:
    b, c, h, w = x.shape
    kh, kw = kernel.shape
    planes = x.reshape(b * c, 1, h, w)
    # zero insertion: sample (i, j) lands at (i*up_y, j*up_x); trailing zeros kept (upfirdn2d.py:156-158)
    grid = planes.new_zeros(b * c, 1, h * up_y, w * up_x)
    grid[:, :, ::up_y, ::up_x] = planes
    # positive pads add zeros, negative pads crop (upfirdn2d.py:160-168)
    grid = F.pad(grid, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y_lo, y_hi = max(-py0, 0), grid.shape[2] - max(-py1, 0)
    x_lo, x_hi = max(-px0, 0), grid.shape[3] - max(-px1, 0)
    grid = grid[:, :, y_lo:y_hi, x_lo:x_hi]
    taps = torch.flip(kernel, [0, 1]).reshape(1, 1, kh, kw).to(x.dtype)   # upfirdn2d.py:174
    full = F.conv2d(grid, taps)
    out = full[:, :, ::down_y, ::down_x]                                  # upfirdn2d.py:184
    return out.reshape(b, c, out.shape[2], out.shape[3])


def upfirdn2d_out_size(n: int, up: int, down: int, p0: int, p1: int, k: int) -> int:
    """op/upfirdn2d.py:100-101."""
    return (n * up + p0 + p1 - k) // down + 1


# --------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------
def pixel_norm(z: Tensor) -> Tensor:
    """model.py:14-15."""
    return z * torch.rsqrt((z * z).mean(dim=1, keepdim=True) + 1e-8)


def equal_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], lr_mul: float = 1.0,
                 activation: bool = False) -> Tensor:
    """model.py:144-157: scale = lr_mul / sqrt(in_dim); bias * lr_mul; optional fused lrelu."""
    scale = lr_mul / math.sqrt(weight.shape[1])
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, None if bias is None else bias * lr_mul)


def mapping_network(sd: Dict[str, Tensor], z: Tensor, n_mlp: int, lr_mlp: float = 0.01) -> Tensor:
    """model.py:376-387: PixelNorm then n_mlp x EqualLinear(lr_mul=lr_mlp, fused_lrelu); the
    Sequential indices start at 1 because slot 0 is the parameter-free PixelNorm."""
    w = pixel_norm(z)
    for i in range(1, n_mlp + 1):
        w = equal_linear(w, sd[f"style.{i}.weight"], sd[f"style.{i}.bias"], lr_mlp, True)
    return w


def modulated_conv2d(x: Tensor, w_latent: Tensor, weight: Tensor, mod_weight: Tensor,
                     mod_bias: Tensor, demodulate: bool = True, upsample: bool = False,
                     downsample: bool = False, blur_taps: Optional[Tensor] = None,
                     blur_pad: Tuple[int, int] = (0, 0)) -> Tensor:
    """model.py:232-273.  weight [1,Cout,Cin,k,k]; per-sample weights are materialised exactly as
    the reference does and fed to a grouped conv with groups = batch."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear(w_latent, mod_weight, mod_bias)                       # :235 (bias_init = 1)
    conv_scale = 1.0 / math.sqrt(cin * k * k)                              # :214-215
    wmod = conv_scale * weight * s.reshape(b, 1, cin, 1, 1)                # :236
    if demodulate:
        d = torch.rsqrt((wmod * wmod).sum(dim=[2, 3, 4]) + 1e-8)           # :239
        wmod = wmod * d.reshape(b, cout, 1, 1, 1)
    if upsample:                                                           # :246-257
        wt = wmod.transpose(1, 2).reshape(b * cin, cout, k, k)
        y = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt, padding=0, stride=2, groups=b)
        y = y.reshape(b, cout, y.shape[2], y.shape[3])
        return upfirdn2d(y, blur_taps, pad=blur_pad)
    if downsample:                                                         # :259-265
        xb = upfirdn2d(x, blur_taps, pad=blur_pad)
        y = F.conv2d(xb.reshape(1, b * cin, xb.shape[2], xb.shape[3]),
                     wmod.reshape(b * cout, cin, k, k), padding=0, stride=2, groups=b)
        return y.reshape(b, cout, y.shape[2], y.shape[3])
    y = F.conv2d(x.reshape(1, b * cin, h, w), wmod.reshape(b * cout, cin, k, k),
                 padding=k // 2, groups=b)                                 # :267-271
    return y.reshape(b, cout, y.shape[2], y.shape[3])


def upconv_blur_pad(n_taps: int = 4, kernel_size: int = 3, factor: int = 2) -> Tuple[int, int]:
    """model.py:198-204."""
    p = (n_taps - factor) - (kernel_size - 1)
    return (p + 1) // 2 + factor - 1, p // 2 + 1


def downconv_blur_pad(n_taps: int = 4, kernel_size: int = 3, factor: int = 2) -> Tuple[int, int]:
    """model.py:206-212."""
    p = (n_taps - factor) + (kernel_size - 1)
    return (p + 1) // 2, p // 2


def skip_upsample_pad(n_taps: int = 4, factor: int = 2) -> Tuple[int, int]:
    """model.py:37-42."""
    p = n_taps - factor
    return (p + 1) // 2 + factor - 1, p // 2


def styled_conv(sd: Dict[str, Tensor], prefix: str, x: Tensor, w_latent: Tensor,
                noise: Optional[Tensor], upsample: bool, blur: Sequence[float],
                gen: Optional[torch.Generator] = None) -> Tensor:
    """model.py:331-337: conv -> + noise_weight * noise -> fused bias + lrelu(0.2) * sqrt(2)."""
    taps = pad = None
    if upsample:
        taps = fir_kernel_2d(blur).to(x.dtype) * 4                         # :77-78
        pad = upconv_blur_pad(len(blur))
    y = modulated_conv2d(x, w_latent, sd[prefix + ".conv.weight"],
                         sd[prefix + ".conv.modulation.weight"], sd[prefix + ".conv.modulation.bias"],
                         True, upsample, False, taps, pad or (0, 0))
    if noise is None:                                                      # :283-285
        noise = torch.randn(y.shape[0], 1, y.shape[2], y.shape[3], generator=gen).to(y.dtype)
    y = y + sd[prefix + ".noise.weight"] * noise                           # :287
    return fused_leaky_relu(y, sd[prefix + ".activate.bias"])              # :335


def to_rgb(sd: Dict[str, Tensor], prefix: str, x: Tensor, w_latent: Tensor,
           skip: Optional[Tensor], blur: Sequence[float]) -> Tensor:
    """model.py:350-359: 1x1 modulated conv without demodulation, + bias, + upsampled skip."""
    y = modulated_conv2d(x, w_latent, sd[prefix + ".conv.weight"],
                         sd[prefix + ".conv.modulation.weight"], sd[prefix + ".conv.modulation.bias"],
                         demodulate=False)
    y = y + sd[prefix + ".bias"]
    if skip is not None:
        taps = fir_kernel_2d(blur).to(x.dtype) * 4                         # :34
        y = y + upfirdn2d(skip, taps, up=2, down=1, pad=skip_upsample_pad(len(blur)))
    return y


def generator_forward(sd: Dict[str, Tensor], size: int, styles: Sequence[Tensor], *,
                      n_mlp: int = 8, lr_mlp: float = 0.01, blur: Sequence[float] = (1, 3, 3, 1),
                      return_latents: bool = False, return_features: bool = False,
                      inject_index: Optional[int] = None, truncation: float = 1.0,
                      truncation_latent: Optional[Tensor] = None, input_is_latent: bool = False,
                      noise: Optional[List[Optional[Tensor]]] = None, randomize_noise: bool = True,
                      gen: Optional[torch.Generator] = None):
    """model.py:470-542, same argument meaning; always returns a 2-tuple."""
    log_size = int(math.log2(size))
    n_latent = 2 * log_size - 2
    num_layers = 2 * (log_size - 2) + 1
    styles = list(styles)
    if not input_is_latent:                                                # :482-483
        styles = [mapping_network(sd, s, n_mlp, lr_mlp) for s in styles]
    if noise is None:                                                      # :485-491
        if randomize_noise:
            noise = [None] * num_layers
        else:
            noise = [sd[f"noises.noise_{i}"] for i in range(num_layers)]
    if truncation < 1:                                                     # :493-501
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    if len(styles) < 2:                                                    # :503-509
        if styles[0].ndim < 3:
            latent = styles[0].unsqueeze(1).repeat(1, n_latent, 1)
        else:
            latent = styles[0]
    else:                                                                  # :511-518
        if inject_index is None:
            raise ValueError("oracle needs an explicit inject_index for style mixing")
        latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                            styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)], 1)
    batch = latent.shape[0]
    out = sd["input.input"].repeat(batch, 1, 1, 1)                         # :296-300
    out = styled_conv(sd, "conv1", out, latent[:, 0], noise[0], False, blur, gen)   # :521
    skip = to_rgb(sd, "to_rgb1", out, latent[:, 1], None, blur)            # :523
    i = 1
    for j in range(log_size - 2):                                          # :526-533
        out = styled_conv(sd, f"convs.{2 * j}", out, latent[:, i], noise[1 + 2 * j], True, blur, gen)
        out = styled_conv(sd, f"convs.{2 * j + 1}", out, latent[:, i + 1], noise[2 + 2 * j], False, blur, gen)
        skip = to_rgb(sd, f"to_rgbs.{j}", out, latent[:, i + 2], skip, blur)
        i += 2
    if return_latents:
        return skip, latent
    if return_features:
        return skip, out
    return skip, None


# --------------------------------------------------------------------------------------------
# deterministic synthetic parameters
# --------------------------------------------------------------------------------------------
def state_dict_spec(size: int, style_dim: int = 512, n_mlp: int = 8,
                    channel_multiplier: int = 2, n_taps: int = 4) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list of Generator.state_dict() as model.py:362-446 registers it."""
    ch = channel_table(channel_multiplier)
    log_size = int(math.log2(size))
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    for i in range(1, n_mlp + 1):
        spec += [(f"style.{i}.weight", (style_dim, style_dim)), (f"style.{i}.bias", (style_dim,))]
    spec.append(("input.input", (1, ch[4], 4, 4)))

    def styled(prefix, cin, cout, up):
        s = [(prefix + ".conv.weight", (1, cout, cin, 3, 3))]
        if up:
            s.append((prefix + ".conv.blur.kernel", (n_taps, n_taps)))
        s += [(prefix + ".conv.modulation.weight", (cin, style_dim)),
              (prefix + ".conv.modulation.bias", (cin,)),
              (prefix + ".noise.weight", (1,)),
              (prefix + ".activate.bias", (cout,))]
        return s

    def rgb(prefix, cin, up):
        s = [(prefix + ".bias", (1, 3, 1, 1))]
        if up:
            s.append((prefix + ".upsample.kernel", (n_taps, n_taps)))
        s += [(prefix + ".conv.weight", (1, 3, cin, 1, 1)),
              (prefix + ".conv.modulation.weight", (cin, style_dim)),
              (prefix + ".conv.modulation.bias", (cin,))]
        return s

    spec += styled("conv1", ch[4], ch[4], False)
    spec += rgb("to_rgb1", ch[4], False)
    conv_items, rgb_items = [], []
    cin = ch[4]
    for i in range(3, log_size + 1):
        cout = ch[2 ** i]
        j = i - 3
        conv_items += styled(f"convs.{2 * j}", cin, cout, True)
        conv_items += styled(f"convs.{2 * j + 1}", cout, cout, False)
        rgb_items += rgb(f"to_rgbs.{j}", cout, True)
        cin = cout
    spec += conv_items + rgb_items
    for layer in range(2 * (log_size - 2) + 1):
        r = 2 ** ((layer + 5) // 2)
        spec.append((f"noises.noise_{layer}", (1, 1, r, r)))
    return spec


def _named_randn(name: str, shape: Sequence[int], seed: int) -> Tensor:
    h = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")
    g = torch.Generator(device="cpu")
    g.manual_seed(h)
    return torch.randn(*shape, generator=g, dtype=torch.float32)


def named_randn(name: str, shape: Sequence[int], seed: int = 0) -> Tensor:
    """Seeded N(0,1) tensor keyed by a string, independent of construction order."""
    return _named_randn(name, shape, seed)


def init_state_dict(size: int, style_dim: int = 512, n_mlp: int = 8, channel_multiplier: int = 2,
                    seed: int = 0, lr_mlp: float = 0.01, blur: Sequence[float] = (1, 3, 3, 1),
                    perturb: float = 0.1) -> Dict[str, Tensor]:
    """Synthetic weights with the reference's init distributions (model.py:134-139,226-228,280,294,
    348; op/fused_act.py:76), each tensor drawn from its own name-keyed stream so both
    implementations can be loaded with identical values.  Parameters the reference initialises to
    zero (noise weights, activation biases, ToRGB biases) are perturbed by N(0, perturb^2) so that
    parity tests exercise them."""
    sd: Dict[str, Tensor] = {}
    k2 = fir_kernel_2d(blur)
    for key, shape in state_dict_spec(size, style_dim, n_mlp, channel_multiplier, len(blur)):
        if key.endswith("blur.kernel") or key.endswith("upsample.kernel"):
            sd[key] = (k2 * 4).clone()
        elif key.startswith("style.") and key.endswith(".weight"):
            sd[key] = _named_randn(key, shape, seed) / lr_mlp
        elif key.startswith("style.") and key.endswith(".bias"):
            sd[key] = perturb * _named_randn(key, shape, seed) / lr_mlp if perturb else torch.zeros(shape)
        elif key.endswith("modulation.bias"):
            sd[key] = torch.ones(shape)
        elif key.endswith("noise.weight") or key.endswith("activate.bias") or (
                key.endswith(".bias") and "to_rgb" in key):
            sd[key] = perturb * _named_randn(key, shape, seed) if perturb else torch.zeros(shape)
        else:  # conv / modulation weights, constant input, noise buffers
            sd[key] = _named_randn(key, shape, seed)
    return sd


def flops_per_image(size: int, channel_multiplier: int = 2) -> float:
    """2*MAC of every modulated conv, up-convs counted in the reference's conv_transpose
    formulation: 90.24 GFLOP @256, 148.52 GFLOP @1024."""
    ch = channel_table(channel_multiplier)
    log_size = int(math.log2(size))
    total = 2.0 * 9 * ch[4] * ch[4] * 16 + 2.0 * 3 * ch[4] * 16
    cin = ch[4]
    for i in range(3, log_size + 1):
        cout, r = ch[2 ** i], 2 ** i
        total += 2.0 * 9 * cin * cout * (r // 2) ** 2      # transposed conv, per input pixel
        total += 2.0 * 9 * cout * cout * r * r
        total += 2.0 * 3 * cout * r * r
        cin = cout
    return total
