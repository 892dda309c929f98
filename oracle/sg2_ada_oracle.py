"""CPU oracle of the StyleGAN2-ADA decoder variant -- TEST INFRASTRUCTURE ONLY.

A functional (state_dict in, tensors out) restatement in plain torch-CPU ops of
`/root/reference/restyle-encoder/models/stylegan2_ada/{generator,utils}.py`, each function citing the lines it
follows.  Only tests/ and tools that CHECK the product may import it; the product never does.

Pinning: the reference holds no tests or golden vectors for this path ("parity unpinned" by the reference's own
tests); `tests/golden/make_golden_ada.py` runs the UNMODIFIED reference module on CPU (it is pure PyTorch) on
weights from `init_state_dict` and commits the outputs (tests/golden/ada.npz); tests/test_oracle_golden.py checks
this restatement against them.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from .sg2_oracle import named_randn


def block_resolutions(img_resolution: int) -> List[int]:
    """generator.py:62-64."""
    return [2 ** i for i in range(2, int(math.log2(img_resolution)) + 1)]


def channels(res: int, channel_base: int = 16384, channel_max: int = 512) -> int:
    """generator.py:66."""
    return min(channel_base // res, channel_max)


def num_ws(img_resolution: int) -> int:
    """generator.py:65."""
    return 2 * (len(block_resolutions(img_resolution)) + 1)


def state_dict_spec(img_resolution: int, z_dim: int = 512, w_dim: int = 512, num_layers: int = 8,
                    img_channels: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    """Keys and shapes of Generator.state_dict() in registration order (generator.py:8-17,58-76,93-103,116-129,
    144-149,174-189,244-262; utils.py:36-44,78-87)."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def layer(prefix, cin, cout, res, up):
        if up:
            spec.append((prefix + ".resampler.kernel", (1, 1, 4, 4)))
        spec.extend([(prefix + ".weight", (cout, cin, 3, 3)), (prefix + ".noise_strength", (1,)), (prefix + ".bias", (cout,)),
                     (prefix + ".noise_const", (res, res)), (prefix + ".affine.weight", (cin, w_dim)),
                     (prefix + ".affine.bias", (cin,))])

    def torgb(prefix, cin):
        spec.extend([(prefix + ".weight", (img_channels, cin, 1, 1)), (prefix + ".bias", (img_channels,)),
                     (prefix + ".affine.weight", (cin, w_dim)), (prefix + ".affine.bias", (cin,))])

    res_list = block_resolutions(img_resolution)
    c0 = channels(res_list[0])
    spec.append(("synthesis.first_block.const", (c0, res_list[0], res_list[0])))
    layer("synthesis.first_block.conv1", c0, c0, res_list[0], False)
    torgb("synthesis.first_block.torgb", c0)
    for i, res in enumerate(res_list[1:]):
        cin, cout = channels(res // 2), channels(res)
        p = f"synthesis.blocks.{i}"
        spec.append((p + ".resampler.kernel", (1, 1, 4, 4)))
        layer(p + ".conv0", cin, cout, res, True)
        layer(p + ".conv1", cout, cout, res, False)
        torgb(p + ".torgb", cout)
    spec.append(("mapping.w_avg", (w_dim,)))
    for i in range(num_layers):
        spec.extend([(f"mapping.layers.{i}.weight", (w_dim, z_dim if i == 0 else w_dim)), (f"mapping.layers.{i}.bias", (w_dim,))])
    return spec


def smooth_kernel() -> Tensor:
    """utils.py:80-86."""
    k = torch.tensor([[1., 3., 3., 1.], [3., 9., 9., 3.], [3., 9., 9., 3.], [1., 3., 3., 1.]])
    return (k / k.sum()).reshape(1, 1, 4, 4)


def init_state_dict(img_resolution: int, z_dim: int = 512, w_dim: int = 512, num_layers: int = 8, seed: int = 0,
                    lr_multiplier: float = 0.01, perturb: float = 0.1) -> Dict[str, Tensor]:
    """Synthetic weights with the reference's init distributions (utils.py:40-41; generator.py:101,146-147,181-186),
    each tensor from its own name-keyed stream.  Parameters the reference initialises to zero (noise strengths,
    biases, w_avg) are perturbed by N(0, perturb^2) so that parity tests exercise them."""
    sd: Dict[str, Tensor] = {}
    for key, shape in state_dict_spec(img_resolution, z_dim, w_dim, num_layers):
        r = named_randn("ada:" + key, shape, seed)
        if key.endswith("resampler.kernel"):
            sd[key] = smooth_kernel()
        elif key.startswith("mapping.layers") and key.endswith(".weight"):
            sd[key] = r / lr_multiplier
        elif key.startswith("mapping.layers") and key.endswith(".bias"):
            sd[key] = perturb * r / lr_multiplier
        elif key.endswith("affine.bias"):
            sd[key] = torch.ones(shape)
        elif key.endswith("noise_strength") or key.endswith(".bias") or key == "mapping.w_avg":
            sd[key] = perturb * r
        else:
            sd[key] = r
    return sd


# ---- ops --------------------------------------------------------------------------------------------------
def fully_connected(x: Tensor, weight: Tensor, bias: Optional[Tensor], lr_multiplier: float = 1.0, lrelu: bool = False) -> Tensor:
    """utils.py:46-52: act(addmm(b * bias_gain, x, (w * weight_gain)^T)) * act_gain."""
    w = weight * (lr_multiplier / math.sqrt(weight.shape[1]))
    b = bias * lr_multiplier if lr_multiplier != 1 else bias
    y = torch.addmm(b.unsqueeze(0), x, w.t())
    return F.leaky_relu(y, 0.2) * math.sqrt(2) if lrelu else y


def smooth_upsample(x: Tensor, kernel: Tensor) -> Tensor:
    """utils.py:89-95: nearest x2 -> replication pad (2,1,2,1) -> conv2d with the 4x4 kernel."""
    b, c, h, w = x.shape
    y = F.interpolate(x.reshape(-1, 1, h, w), scale_factor=2, mode="nearest")
    y = F.pad(y, (2, 1, 2, 1), mode="replicate")
    return F.conv2d(y, kernel.to(x.dtype)).view(b, c, h * 2, w * 2)


def modulated_conv2d(x: Tensor, weight: Tensor, styles: Tensor, padding: int = 0, demodulate: bool = True) -> Tensor:
    """utils.py:118-137."""
    n = x.shape[0]
    o, i, kh, kw = weight.shape
    w = weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        w = w * d.reshape(n, -1, 1, 1, 1)
    y = F.conv2d(x.reshape(1, -1, *x.shape[2:]), w.reshape(-1, i, kh, kw), padding=padding, groups=n)
    return y.reshape(n, -1, *y.shape[2:])


def clamp_gain(x: Tensor, g: float, c: float) -> Tensor:
    """utils.py:6-7."""
    return torch.clamp(x * g, -c, c)


def synthesis_layer(sd, p: str, x: Tensor, w: Tensor, noise: Optional[Tensor], up: bool) -> Tensor:
    """SynthesisLayer2.forward, generator.py:186-204.  noise: None ('none'), or the tensor to add before scaling."""
    styles = fully_connected(w, sd[p + ".affine.weight"], sd[p + ".affine.bias"])
    y = modulated_conv2d(x, sd[p + ".weight"], styles, padding=1)
    if up:
        y = smooth_upsample(y, sd[p + ".resampler.kernel"])
    if noise is not None:
        y = y + noise * sd[p + ".noise_strength"]
    return clamp_gain(F.leaky_relu(y + sd[p + ".bias"][None, :, None, None], 0.2), math.sqrt(2), 256.0)


def torgb_layer(sd, p: str, x: Tensor, w: Tensor) -> Tensor:
    """ToRGBLayer2.forward, generator.py:148-151."""
    cin = sd[p + ".weight"].shape[1]
    styles = fully_connected(w, sd[p + ".affine.weight"], sd[p + ".affine.bias"]) * (1 / math.sqrt(cin))
    y = modulated_conv2d(x, sd[p + ".weight"], styles, demodulate=False)
    return torch.clamp(y + sd[p + ".bias"][None, :, None, None], -256, 256)


def mapping_network(sd, z: Tensor, n_ws: int, num_layers: int, truncation_psi: float = 1.0,
                    truncation_cutoff: Optional[int] = None, lr_multiplier: float = 0.01) -> Tensor:
    """MappingNetwork.forward in eval mode, generator.py:264-286 (normalize_2nd_moment: utils.py:10-11)."""
    x = z * (z.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    for i in range(num_layers):
        x = fully_connected(x, sd[f"mapping.layers.{i}.weight"], sd[f"mapping.layers.{i}.bias"], lr_multiplier, lrelu=True)
    x = x.unsqueeze(1).repeat([1, n_ws, 1])
    if truncation_psi != 1:
        if truncation_cutoff is None:
            x = sd["mapping.w_avg"].lerp(x, truncation_psi)
        else:
            x[:, :truncation_cutoff] = sd["mapping.w_avg"].lerp(x[:, :truncation_cutoff], truncation_psi)
    return x


def synthesis_network(sd, img_resolution: int, ws: Tensor, noise_mode: str = "const",
                      noises: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """SynthesisNetwork.forward + SynthesisPrologue / SynthesisBlock.forward, generator.py:78-88,105-111,131-139.
    noise_mode 'const' uses the noise_const buffers; `noises` (layer prefix -> [B,1,r,r]) overrides them."""
    res_list = block_resolutions(img_resolution)
    split = [ws[:, 0:2, :]] + [ws[:, 2 * n + 1: 2 * n + 4, :] for n in range(len(res_list))]

    def nz(p):
        if noises is not None and p in noises:
            return noises[p]
        return sd[p + ".noise_const"] if noise_mode == "const" else None

    b = ws.shape[0]
    x = sd["synthesis.first_block.const"].unsqueeze(0).repeat(b, 1, 1, 1)
    p = "synthesis.first_block"
    x = synthesis_layer(sd, p + ".conv1", x, split[0][:, 0], nz(p + ".conv1"), False)
    img = torgb_layer(sd, p + ".torgb", x, split[0][:, 1])
    for i in range(len(res_list) - 1):
        p = f"synthesis.blocks.{i}"
        w3 = split[i + 1]
        x = synthesis_layer(sd, p + ".conv0", x, w3[:, 0], nz(p + ".conv0"), True)
        x = synthesis_layer(sd, p + ".conv1", x, w3[:, 1], nz(p + ".conv1"), False)
        y = torgb_layer(sd, p + ".torgb", x, w3[:, 2])
        img = smooth_upsample(img, sd[p + ".resampler.kernel"]) + y
    return img


def generator_forward(sd, img_resolution: int, z: Tensor, num_layers: int = 8, input_is_latent: bool = False,
                      truncation_psi: float = 1.0, truncation_cutoff: Optional[int] = None) -> Tensor:
    """Generator.forward with randomize_noise falsy (noise_mode 'const'), generator.py:19-40."""
    ws = z if input_is_latent else mapping_network(sd, z, num_ws(img_resolution), num_layers, truncation_psi, truncation_cutoff)
    return synthesis_network(sd, img_resolution, ws, "const")
