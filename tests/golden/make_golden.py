#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Runs only in the authoring container (needs /root/reference; the GPU box does not have it).
The reference's model.py is imported untouched; its three op names are rebound to the reference's
own PyTorch statements of the ops exactly as BASELINE.md section 3 / SURVEY.md section 8c describe
(`upfirdn2d_native` from op/upfirdn2d.py:150-184 after injecting the missing `F` import, and the
leaky-relu formula of fused_bias_act_kernel.cu:26-47).  The JIT build of the CUDA extensions is
skipped by stubbing `torch.utils.cpp_extension.load` (no GPU here, they could not run anyway).

Outputs (committed): tests/golden/ops.npz, layers.npz, generator.npz, state_dict_keys.json

    python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import sg2_oracle as O  # noqa: E402  (used for seeded inputs + a cross-check print)


def import_reference():
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: types.SimpleNamespace()
    sys.path.insert(0, "/root/reference")
    import backbone.stylegan2.op  # noqa: F401
    um = sys.modules["backbone.stylegan2.op.upfirdn2d"]
    um.F = F
    import backbone.stylegan2.model as M

    def upfirdn2d_cpu(x, k, up=1, down=1, pad=(0, 0)):
        b, c, h, w = x.shape
        o = um.upfirdn2d_native(x.reshape(-1, h, w, 1), k, up, up, down, down,
                                pad[0], pad[1], pad[0], pad[1])
        return o.view(b, c, o.shape[1], o.shape[2])

    def fused_leaky_relu_cpu(x, b, negative_slope=0.2, scale=2 ** 0.5):
        return scale * F.leaky_relu(x + b.view(1, -1, *[1] * (x.ndim - 2)), negative_slope)

    class FusedLeakyReLUCPU(torch.nn.Module):
        def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
            super().__init__()
            self.bias = torch.nn.Parameter(torch.zeros(channel))
            self.negative_slope, self.scale = negative_slope, scale

        def forward(self, x):
            return fused_leaky_relu_cpu(x, self.bias, self.negative_slope, self.scale)

    M.upfirdn2d, M.fused_leaky_relu, M.FusedLeakyReLU = upfirdn2d_cpu, fused_leaky_relu_cpu, FusedLeakyReLUCPU
    return M, um, upfirdn2d_cpu, fused_leaky_relu_cpu


# ---------------------------------------------------------------------------------------------
# case tables (shared with the tests through the npz "meta" json)
# ---------------------------------------------------------------------------------------------
UPFIRDN_CASES = [
    # name, (B,C,H,W), kernel spec, up, down, pad
    ("blur_upconv", (2, 3, 17, 17), "blur4x4", 1, 1, (1, 1)),      # model.py:198-204 after conv_transpose
    ("skip_up2", (2, 3, 8, 8), "blur4x4", 2, 1, (2, 1)),           # model.py:37-47 ToRGB skip
    ("down2", (2, 3, 16, 16), "blur1", 1, 2, (1, 1)),              # model.py:58-68 Downsample
    ("blur_pad22", (1, 4, 9, 9), "blur1", 1, 1, (2, 2)),           # discriminator blur model.py:558-564
    ("rand3x3_asym", (1, 2, 7, 11), "rand3x3", 1, 1, (1, 1)),      # asymmetric taps: proves the flip
    ("rand4x4_up2", (1, 2, 5, 6), "rand4x4", 2, 1, (2, 1)),
    ("rand4x4_down2", (1, 2, 12, 10), "rand4x4", 1, 2, (1, 1)),
    ("k2_up2", (1, 2, 6, 6), "rand2x2", 2, 1, (1, 0)),
    ("k2_down2", (1, 2, 8, 8), "rand2x2", 1, 2, (0, 0)),
    ("negpad_crop", (1, 2, 10, 10), "rand3x3", 1, 1, (-1, 2)),     # negative pad crops (upfirdn2d.py:163-168)
    ("tiny_4x4", (3, 5, 4, 4), "blur4x4", 2, 1, (2, 1)),
    ("bwd_blur", (2, 3, 16, 16), "blur4x4", 1, 1, (2, 2)),         # backward of blur_upconv (upfirdn2d.py:108-113)
    ("bwd_skip", (2, 3, 16, 16), "blur4x4", 1, 2, (1, 1)),         # backward of skip_up2
]


def fir(spec):
    if spec == "blur4x4":
        return O.fir_kernel_2d([1, 3, 3, 1]) * 4
    if spec == "blur1":
        return O.fir_kernel_2d([1, 3, 3, 1])
    n = int(spec[4])
    return O.named_randn("fir:" + spec, (n, n), 7)


LRELU_CASES = [("mlp_2d", (5, 37)), ("act_4d", (2, 6, 5, 7)), ("act_3d", (2, 4, 9)), ("one_px", (3, 8, 1, 1))]

LAYER_CASES = [
    # name, cin, cout, k, style_dim, demod, up, down, (B,H,W)
    ("plain3", 8, 12, 3, 16, True, False, False, (2, 6, 6)),
    ("up3", 8, 12, 3, 16, True, True, False, (2, 5, 5)),
    ("down3", 8, 6, 3, 16, True, False, True, (2, 8, 8)),
    ("rgb1", 12, 3, 1, 16, False, False, False, (2, 6, 6)),
    ("plain3_odd", 5, 7, 3, 9, True, False, False, (3, 4, 7)),
    ("up3_64", 64, 32, 3, 32, True, True, False, (1, 8, 8)),
]

GEN_CASES = [
    # name, size, n_mlp, channel_multiplier, batch, mode
    ("g16_z", 16, 2, 2, 2, "z"),
    ("g16_wplus_noise", 16, 2, 2, 2, "wplus_noise"),
    ("g16_trunc", 16, 2, 2, 2, "trunc"),
    ("g16_mix", 16, 2, 2, 2, "mix"),
    ("g16_features", 16, 2, 2, 1, "features"),
    ("g64_z", 64, 8, 2, 1, "z"),
    ("g64_cm1_w", 64, 4, 1, 2, "w"),
    ("g256_z", 256, 8, 2, 1, "z"),
    ("g1024_wplus", 1024, 8, 2, 1, "wplus"),
]


def gen_inputs(name, size, n_mlp, batch, mode, sd):
    """Seeded inputs for a generator case; returns (styles list, kwargs)."""
    log_size = int(np.log2(size))
    n_latent, num_layers = 2 * log_size - 2, 2 * (log_size - 2) + 1
    z = O.named_randn(name + ":z", (batch, 512), 1)
    kw = dict(randomize_noise=False)
    if mode == "z":
        return [z], kw
    w = O.mapping_network(sd, z, n_mlp)
    if mode == "w":
        kw["input_is_latent"] = True
        return [w], kw
    if mode in ("wplus", "wplus_noise"):
        wp = w[:, None].repeat(1, n_latent, 1) + 0.1 * O.named_randn(name + ":dw", (batch, n_latent, 512), 1)
        kw["input_is_latent"] = True
        kw["return_latents"] = True
        if mode == "wplus_noise":
            kw["noise"] = [O.named_randn(f"{name}:noise{i}", (batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 1)
                           for i in range(num_layers)]
        return [wp], kw
    if mode == "trunc":
        kw["truncation"] = 0.7
        kw["truncation_latent"] = O.mapping_network(sd, O.named_randn(name + ":zt", (64, 512), 1), n_mlp).mean(0, keepdim=True)
        return [z], kw
    if mode == "mix":
        kw["inject_index"] = 3
        kw["return_latents"] = True
        return [z, O.named_randn(name + ":z2", (batch, 512), 1)], kw
    if mode == "features":
        kw["return_features"] = True
        return [z], kw
    raise ValueError(mode)


def main():
    torch.set_grad_enabled(False)
    M, um, ref_upfirdn2d, ref_lrelu = import_reference()
    out_ops, out_layers, out_gen = {}, {}, {}

    # ---- ops ---------------------------------------------------------------------------------
    for name, shape, kspec, up, down, pad in UPFIRDN_CASES:
        x = O.named_randn("upfirdn:" + name, shape, 3)
        k = fir(kspec)
        y = ref_upfirdn2d(x, k, up, down, pad)
        out_ops["upfirdn2d/" + name] = y.numpy()
        d = (O.upfirdn2d(x, k, up, down, pad) - y).abs().max().item()
        print(f"upfirdn2d {name:16s} out {tuple(y.shape)}  oracle-vs-ref max|d| = {d:.2e}")
    for name, shape in LRELU_CASES:
        x = O.named_randn("lrelu:" + name, shape, 3)
        b = O.named_randn("lrelu_b:" + name, (shape[1],), 3)
        y = ref_lrelu(x, b)
        out_ops["lrelu/" + name] = y.numpy()
        # backward through the reference's formula (autograd of the shim == fused_act.py:20-38 math)
        with torch.enable_grad():
            xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yr = ref_lrelu(xr, br)
            gy = O.named_randn("lrelu_gy:" + name, shape, 3)
            gx, gb = torch.autograd.grad(yr, [xr, br], gy)
        out_ops["lrelu_gx/" + name], out_ops["lrelu_gb/" + name] = gx.numpy(), gb.numpy()
        d = (O.fused_leaky_relu(x, b) - y).abs().max().item()
        print(f"lrelu     {name:16s} oracle-vs-ref max|d| = {d:.2e}")

    # ---- single ModulatedConv2d layers (forward + grads) ---------------------------------------
    for name, cin, cout, k, sdim, demod, up, down, (b, h, w) in LAYER_CASES:
        layer = M.ModulatedConv2d(cin, cout, k, sdim, demodulate=demod, upsample=up, downsample=down)
        wt = O.named_randn(f"layer:{name}:weight", (1, cout, cin, k, k), 5)
        mw = O.named_randn(f"layer:{name}:mod_w", (cin, sdim), 5)
        mb = 1 + 0.1 * O.named_randn(f"layer:{name}:mod_b", (cin,), 5)
        layer.weight.data.copy_(wt), layer.modulation.weight.data.copy_(mw), layer.modulation.bias.data.copy_(mb)
        x = O.named_randn(f"layer:{name}:x", (b, cin, h, w), 5)
        s = O.named_randn(f"layer:{name}:s", (b, sdim), 5)
        with torch.enable_grad():
            xr, sr = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
            y = layer(xr, sr)
            gy = O.named_randn(f"layer:{name}:gy", tuple(y.shape), 5)
            gx, gs, gw, gmw, gmb = torch.autograd.grad(
                y, [xr, sr, layer.weight, layer.modulation.weight, layer.modulation.bias], gy)
        for key, val in (("y", y), ("gx", gx), ("gs", gs), ("gw", gw), ("gmw", gmw), ("gmb", gmb)):
            out_layers[f"{name}/{key}"] = val.detach().numpy()
        taps = pad = None
        if up:
            taps, pad = O.fir_kernel_2d([1, 3, 3, 1]) * 4, O.upconv_blur_pad()
        if down:
            taps, pad = O.fir_kernel_2d([1, 3, 3, 1]), O.downconv_blur_pad()
        yo = O.modulated_conv2d(x, s, wt, mw, mb, demod, up, down, taps, pad or (0, 0))
        print(f"layer     {name:16s} out {tuple(y.shape)}  oracle-vs-ref max|d| = {(yo - y).abs().max().item():.2e}")

    # ---- whole generator ------------------------------------------------------------------------
    keys = {}
    for name, size, n_mlp, cm, batch, mode in GEN_CASES:
        sd = O.init_state_dict(size, 512, n_mlp, cm, seed=0)
        G = M.Generator(size, 512, n_mlp, channel_multiplier=cm)
        ref_keys = [(k, list(v.shape)) for k, v in G.state_dict().items()]
        assert ref_keys == [(k, list(s)) for k, s in O.state_dict_spec(size, 512, n_mlp, cm)], name
        keys[f"{size}_{n_mlp}_{cm}"] = ref_keys
        G.load_state_dict(sd, strict=True)
        G.eval()
        styles, kw = gen_inputs(name, size, n_mlp, batch, mode, sd)
        img, aux = G(styles, **kw)
        img_o, aux_o = O.generator_forward(sd, size, styles, n_mlp=n_mlp, **kw)
        d = (img_o - img).abs().max().item()
        print(f"generator {name:16s} img {tuple(img.shape)} range [{img.min():.2f},{img.max():.2f}] "
              f"oracle-vs-ref max|d| = {d:.2e}")
        if size >= 1024:      # 12 MB/img is too big to commit: keep a strided lattice + moments
            out_gen[name + "/img_lattice8"] = img[:, :, 3::8, 5::8].numpy()
            out_gen[name + "/img_moments"] = np.array(
                [img.double().mean().item(), img.double().std().item(), img.double().abs().mean().item(),
                 img.min().item(), img.max().item()], dtype=np.float64)
        else:
            out_gen[name + "/img"] = img.numpy()
        if aux is not None:
            out_gen[name + "/aux"] = aux.numpy()
        # checksum of the regenerated weights so a drifting RNG is caught, not silently compared
        out_gen[name + "/sd_checksum"] = np.array(
            [float(sum(v.double().abs().sum() for v in sd.values())), float(sd["conv1.conv.weight"][0, 3, 5, 1, 2])],
            dtype=np.float64)

    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out_ops)
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **out_layers)
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out_gen)
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0)
    for fn in ("ops.npz", "layers.npz", "generator.npz", "state_dict_keys.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
