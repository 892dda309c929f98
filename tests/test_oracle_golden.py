"""CPU: the oracle (test infrastructure) is pinned to golden vectors produced by the reference
itself (tests/golden/make_golden.py ran /root/reference's model.py + upfirdn2d_native on CPU)."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_upfirdn2d_oracle_matches_reference(oracle, golden, cases):
    for name, shape, kspec, up, down, pad in cases.UPFIRDN_CASES:
        x = oracle.named_randn("upfirdn:" + name, shape, 3)
        y = oracle.upfirdn2d(x, cases.fir(kspec), up, down, pad)
        ref = torch.from_numpy(golden["ops"]["upfirdn2d/" + name])
        assert y.shape == ref.shape, name
        assert torch.equal(y, ref), name          # same library calls -> bit-exact


def test_lrelu_oracle_matches_reference(oracle, golden, cases):
    for name, shape in cases.LRELU_CASES:
        x = oracle.named_randn("lrelu:" + name, shape, 3)
        b = oracle.named_randn("lrelu_b:" + name, (shape[1],), 3)
        y = oracle.fused_leaky_relu(x, b)
        assert torch.equal(y, torch.from_numpy(golden["ops"]["lrelu/" + name])), name
        gy = oracle.named_randn("lrelu_gy:" + name, shape, 3)
        gx, gb = oracle.fused_leaky_relu_backward(gy, y)
        np.testing.assert_allclose(gx.numpy(), golden["ops"]["lrelu_gx/" + name], rtol=0, atol=1e-6)
        np.testing.assert_allclose(gb.numpy(), golden["ops"]["lrelu_gb/" + name], rtol=1e-5, atol=1e-5)


def _layer_inputs(oracle, name, cin, cout, k, sdim, b, h, w):
    wt = oracle.named_randn(f"layer:{name}:weight", (1, cout, cin, k, k), 5)
    mw = oracle.named_randn(f"layer:{name}:mod_w", (cin, sdim), 5)
    mb = 1 + 0.1 * oracle.named_randn(f"layer:{name}:mod_b", (cin,), 5)
    x = oracle.named_randn(f"layer:{name}:x", (b, cin, h, w), 5)
    s = oracle.named_randn(f"layer:{name}:s", (b, sdim), 5)
    return wt, mw, mb, x, s


def test_modconv_oracle_matches_reference(oracle, golden, cases):
    for name, cin, cout, k, sdim, demod, up, down, (b, h, w) in cases.LAYER_CASES:
        wt, mw, mb, x, s = _layer_inputs(oracle, name, cin, cout, k, sdim, b, h, w)
        taps = pad = None
        if up:
            taps, pad = oracle.fir_kernel_2d([1, 3, 3, 1]) * 4, oracle.upconv_blur_pad()
        if down:
            taps, pad = oracle.fir_kernel_2d([1, 3, 3, 1]), oracle.downconv_blur_pad()
        y = oracle.modulated_conv2d(x, s, wt, mw, mb, demod, up, down, taps, pad or (0, 0))
        assert torch.equal(y, torch.from_numpy(golden["layers"][name + "/y"])), name


def test_generator_oracle_matches_reference(oracle, golden, cases):
    for name, size, n_mlp, cm, batch, mode in cases.GEN_CASES:
        sd = oracle.init_state_dict(size, 512, n_mlp, cm, seed=0)
        chk = golden["generator"][name + "/sd_checksum"]
        total = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(total - chk[0]) <= 1e-9 * abs(chk[0]), "seeded weights drifted (torch RNG changed?)"
        assert float(sd["conv1.conv.weight"][0, 3, 5, 1, 2]) == chk[1]
        styles, kw = cases.gen_inputs(name, size, n_mlp, batch, mode, sd)
        with torch.no_grad():
            img, aux = oracle.generator_forward(sd, size, styles, n_mlp=n_mlp, **kw)
        if size >= 1024:
            ref = torch.from_numpy(golden["generator"][name + "/img_lattice8"])
            assert torch.equal(img[:, :, 3::8, 5::8], ref), name
        else:
            assert torch.equal(img, torch.from_numpy(golden["generator"][name + "/img"])), name
        if aux is not None:
            assert torch.equal(aux, torch.from_numpy(golden["generator"][name + "/aux"])), name


def test_state_dict_spec_matches_reference(oracle):
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    for tag, ref in keys.items():
        size, n_mlp, cm = (int(v) for v in tag.split("_"))
        spec = [(k, list(s)) for k, s in oracle.state_dict_spec(size, 512, n_mlp, cm)]
        assert spec == [(k, list(s)) for k, s in ref], tag
    assert len(oracle.state_dict_spec(256)) == 135      # SURVEY.md section 5 [probe]


def test_flops_table(oracle):
    assert abs(oracle.flops_per_image(256) / 1e9 - 90.24) < 0.2
    assert abs(oracle.flops_per_image(1024) / 1e9 - 148.52) < 0.3


# ---- plain-C oracle (oracle/sg2_oracle_ops.c) against the same golden vectors ----------------------
def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_oracle_upfirdn2d(c_oracle, oracle, golden, cases):
    for name, shape, kspec, up, down, pad in cases.UPFIRDN_CASES:
        x = oracle.named_randn("upfirdn:" + name, shape, 3).numpy()
        k = cases.fir(kspec).numpy().astype(np.float32)
        ref = golden["ops"]["upfirdn2d/" + name]
        out = np.empty_like(ref)
        b, c, h, w = shape
        rc = c_oracle.oracle_upfirdn2d_f32(_fp(out), _fp(np.ascontiguousarray(x)), _fp(k), ctypes.c_int64(b * c),
                                           h, w, 1, k.shape[0], k.shape[1], up, up, down, down,
                                           pad[0], pad[1], pad[0], pad[1])
        assert rc == 0
        np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6, err_msg=name)


def test_c_oracle_bias_act(c_oracle, oracle, golden, cases):
    for name, shape in cases.LRELU_CASES:
        x = oracle.named_randn("lrelu:" + name, shape, 3).numpy()
        b = oracle.named_randn("lrelu_b:" + name, (shape[1],), 3).numpy()
        out = np.empty_like(x)
        step = int(np.prod(shape[2:])) if len(shape) > 2 else 1
        c_oracle.oracle_fused_bias_act_f32(_fp(out), _fp(x), _fp(b), None, ctypes.c_int64(x.size),
                                           ctypes.c_int64(step), ctypes.c_int64(shape[1]), 3, 0,
                                           ctypes.c_float(0.2), ctypes.c_float(2 ** 0.5))
        np.testing.assert_array_equal(out, golden["ops"]["lrelu/" + name])


def test_c_oracle_modconv(c_oracle, oracle, golden, cases):
    for name, cin, cout, k, sdim, demod, up, down, (b, h, w) in cases.LAYER_CASES:
        wt, mw, mb, x, s = _layer_inputs(oracle, name, cin, cout, k, sdim, b, h, w)
        style = oracle.equal_linear(s, mw, mb).numpy()
        xin = x
        if down:
            xin = oracle.upfirdn2d(x, oracle.fir_kernel_2d([1, 3, 3, 1]), pad=oracle.downconv_blur_pad())
        hh, ww = xin.shape[2:]
        mode = 1 if up else (2 if down else 0)
        oh = hh if mode == 0 else ((hh - 1) * 2 + k if mode == 1 else (hh - k) // 2 + 1)
        ow = ww if mode == 0 else ((ww - 1) * 2 + k if mode == 1 else (ww - k) // 2 + 1)
        out = np.empty((b, cout, oh, ow), np.float32)
        rc = c_oracle.oracle_modconv2d_f32(_fp(out), _fp(np.ascontiguousarray(xin.numpy())),
                                           _fp(np.ascontiguousarray(wt[0].numpy())), _fp(style),
                                           b, cin, cout, hh, ww, k, int(demod), mode)
        assert rc == 0
        y = torch.from_numpy(out)
        if up:
            y = oracle.upfirdn2d(y, oracle.fir_kernel_2d([1, 3, 3, 1]) * 4, pad=oracle.upconv_blur_pad())
        np.testing.assert_allclose(y.numpy(), golden["layers"][name + "/y"], rtol=0, atol=2e-5, err_msg=name)


def test_oracle_autograd_matches_reference_gradients():
    """whole-decoder gradients: fp64 autograd through the oracle restatements == fp64 autograd through the unmodified
    reference modules (tests/golden/grads.npz, make_golden_grads.py) -- the `-m gpu` gradient tests compare the CUDA
    path with the oracle's autograd, this pins that autograd to the reference"""
    import make_golden_grads as MG
    from oracle import sg2_ada_oracle as A
    from oracle import sg2_oracle as O
    g = np.load(os.path.join(GOLDEN, "grads.npz"))
    for name, size, n_mlp, batch in MG.ROS_CASES:
        sd = {k: v.double() for k, v in O.init_state_dict(size, 512, n_mlp, 2, seed=0).items()}
        lat, noise, gy = MG.ros_inputs(name, size, batch)
        ld = lat.double().requires_grad_(True)
        nd = [n.double().requires_grad_(True) for n in noise]
        img, _ = O.generator_forward(sd, size, [ld], n_mlp=n_mlp, input_is_latent=True, noise=nd)
        grads = torch.autograd.grad(img, [ld] + nd, gy.double())
        ref_img = torch.from_numpy(g[name + "/image"])
        assert (img.detach() - ref_img).abs().max() <= 1e-10 * ref_img.abs().max(), name
        refs = [g[name + "/g_latent"]] + [g[f"{name}/g_noise{i}"] for i in range(len(nd))]
        for i, (a, r) in enumerate(zip(grads, refs)):
            r = torch.from_numpy(r)
            assert a.shape == r.shape and (a - r).abs().max() <= 1e-9 * max(1.0, r.abs().max().item()), (name, i)
    for name, res, nl, batch in MG.ADA_CASES:
        sd = {k: v.double() for k, v in A.init_state_dict(res, 512, 512, nl, seed=0).items()}
        ws, gy = MG.ada_inputs(name, res, batch)
        wd = ws.double().requires_grad_(True)
        img = A.synthesis_network(sd, res, wd, "const")
        (ga,) = torch.autograd.grad(img, [wd], gy.double())
        ref_img, r = torch.from_numpy(g[name + "/image"]), torch.from_numpy(g[name + "/g_ws"])
        assert (img.detach() - ref_img).abs().max() <= 1e-10 * ref_img.abs().max(), name
        assert ga.shape == r.shape and (ga - r).abs().max() <= 1e-9 * max(1.0, r.abs().max().item()), name
