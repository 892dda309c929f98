"""GPU: the ADA decoder variant on the sg2_b200 kernels vs the golden vectors of the unmodified reference."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ada():
    return importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator"), \
        importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.utils")


def test_smooth_upsample_golden(sg2):
    from oracle.sg2_oracle import named_randn
    gen, U = _ada()
    g = np.load(os.path.join(GOLDEN, "ada.npz"))
    up = U.SmoothUpsample().to(DEV)
    for name, shape in (("up_5x7", (2, 3, 5, 7)), ("up_1x1", (1, 2, 1, 1)), ("up_16", (1, 4, 16, 16))):
        x = named_randn("ada:up:" + name, shape, 2)
        y = up(x.to(DEV)).cpu()
        np.testing.assert_allclose(y.numpy(), g["smooth_upsample/" + name], rtol=0, atol=2e-6, err_msg=name)
    # low precision storage, fp32 math
    x = torch.randn(2, 3, 9, 11)
    from oracle import sg2_ada_oracle as A
    for dt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)):
        y = up(x.to(DEV).to(dt)).float().cpu()
        ref = A.smooth_upsample(x.to(dt).float(), A.smooth_kernel())
        assert y.shape == ref.shape and (y - ref).abs().max() <= tol * ref.abs().max()


def test_fused_epilogues_vs_oracle(sg2):
    from oracle import sg2_ada_oracle as A
    gen, U = _ada()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 6, 7, generator=g)
    noise = torch.randn(2, 1, 12, 14, generator=g)
    ns, bias = torch.tensor([0.3]), torch.randn(5, generator=g)
    k = A.smooth_kernel()
    ref = A.clamp_gain(torch.nn.functional.leaky_relu(A.smooth_upsample(x, k) + noise * ns + bias[None, :, None, None], 0.2),
                       2 ** 0.5, 1.5)                                           # clamp low enough to bite
    y = U.smooth_upsample2x(x.to(DEV), k.to(DEV), noise.to(DEV), ns.to(DEV), bias.to(DEV), None, act=3, gain=2 ** 0.5,
                            clamp=1.5).cpu()
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=3e-6)
    add = torch.randn(2, 5, 12, 14, generator=g)
    y = U.smooth_upsample2x(x.to(DEV), k.to(DEV), addend=add.to(DEV)).cpu()
    np.testing.assert_allclose(y.numpy(), (A.smooth_upsample(x, k) + add).numpy(), rtol=0, atol=3e-6)
    n1 = torch.randn(1, 1, 6, 7, generator=g)                                     # one map broadcast over the batch
    y = U.ada_bias_act(x.to(DEV), n1.to(DEV), ns.to(DEV), bias.to(DEV), act=3, gain=2 ** 0.5, clamp=2.0).cpu()
    ref = A.clamp_gain(torch.nn.functional.leaky_relu(x + n1 * ns + bias[None, :, None, None], 0.2), 2 ** 0.5, 2.0)
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=2e-6)


def test_ada_generator_golden(sg2):
    import make_golden_ada as M
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    gen, U = _ada()
    g = np.load(os.path.join(GOLDEN, "ada.npz"))
    for name, res, nl, b, mode, (psi, cutoff) in M.CASES:
        sd = A.init_state_dict(res, 512, 512, nl, seed=0)
        G = gen.Generator(512, 512, nl, res, 3)
        G.load_state_dict(sd, strict=True)
        G = G.to(DEV).eval()
        ref = torch.from_numpy(g[name + "/image"])
        with torch.no_grad():
            if mode == "z":
                z = named_randn("ada:z:" + name, (b, 512), 1).to(DEV)
                ws = G.mapping(z, truncation_psi=psi, truncation_cutoff=cutoff)
                wref = torch.from_numpy(g[name + "/ws"])
                assert (ws.cpu() - wref).abs().max() <= 2e-4 * wref.abs().max(), name
                img = G.synthesis(ws, "const")[0]
                if psi == 1:
                    img2, none = G([z], randomize_noise=False)
                    assert none is None and torch.equal(img, img2)
            else:
                ws = named_randn("ada:w:" + name, (b, A.num_ws(res), 512), 1).to(DEV)
                img, lat = G([ws], input_is_latent=True, randomize_noise=False, return_latents=True)
                assert lat is ws
        d = (img.cpu() - ref).abs().max().item()
        assert img.shape == ref.shape and d <= 1e-3, (name, d)           # north_star: fp32 within 1e-3 max-abs
        assert d <= 2e-4 * ref.abs().max().item() + 1e-5, (name, d)


def test_ada_generator_random_noise_and_mapping_grad(sg2):
    gen, U = _ada()
    G = gen.Generator(512, 512, 2, 32, 3).to(DEV).eval()
    z = torch.randn(2, 512, device=DEV)
    with torch.no_grad():
        for m in G.modules():
            if hasattr(m, "noise_strength"):
                m.noise_strength.fill_(0.5)
        a, _ = G([z], randomize_noise=True)
        b, _ = G([z], randomize_noise=True)
        c, _ = G([z], randomize_noise=False)
        d, _ = G([z], randomize_noise=False)
    assert not torch.equal(a, b) and torch.equal(c, d) and a.shape == (2, 3, 32, 32)
    assert G.mean_latent(64).shape == (1, G.num_ws, 512) and G.get_latent(z).shape == (2, G.num_ws, 512)
    w = G.get_latent(z)
    assert w.requires_grad                                   # the mapping network keeps an autograd path
    with torch.no_grad():
        assert (w - G.get_latent(z)).abs().max() <= 2e-4 * w.abs().max()
    with pytest.raises(RuntimeError, match="inference-only"):                     # the fused epilogues themselves have no autograd
        U.ada_bias_act(c.requires_grad_(True), None, None, None)


def test_smooth_upsample_backward_is_the_adjoint(sg2):
    from oracle import sg2_ada_oracle as A
    gen, U = _ada()
    k = A.smooth_kernel()
    for shape in ((2, 3, 5, 7), (1, 2, 1, 1), (1, 1, 2, 3), (2, 2, 16, 16), (1, 1, 33, 65)):
        g = torch.Generator().manual_seed(sum(shape))
        x = torch.randn(shape, generator=g, dtype=torch.float64).requires_grad_(True)
        gy = torch.randn(shape[0], shape[1], 2 * shape[2], 2 * shape[3], generator=g)
        A.smooth_upsample(x, k.double()).backward(gy.double())
        xd = x.detach().float().to(DEV).requires_grad_(True)
        y = U.SmoothUpsampleFunction.apply(xd, k.to(DEV))
        y.backward(gy.to(DEV))
        np.testing.assert_allclose(xd.grad.cpu().numpy(), x.grad.float().numpy(), rtol=0, atol=3e-6, err_msg=str(shape))
    x = torch.randn(2, 3, 9, 11)
    gy = torch.randn(2, 3, 18, 22)
    xr = x.clone().requires_grad_(True)
    A.smooth_upsample(xr, k).backward(gy)
    for dt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)):
        xd = x.to(DEV).to(dt).requires_grad_(True)
        U.SmoothUpsampleFunction.apply(xd, k.to(DEV)).backward(gy.to(DEV).to(dt))
        ref = torch.autograd.grad(A.smooth_upsample(xr, k), xr, gy.to(dt).float())[0]
        assert xd.grad.dtype == dt and (xd.grad.float().cpu() - ref).abs().max() <= tol * ref.abs().max()


@pytest.mark.parametrize("res,b,nl", [(32, 2, 2), (16, 3, 2)])
def test_ada_decoder_gradients_vs_oracle(sg2, res, b, nl):
    """forward + backward through the frozen and the trainable decoder (the ReStyle coaches back-propagate an image
    loss to the encoder through it, coach_restyle_psp.py:86-101) vs torch-CPU autograd of the oracle."""
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    gen, U = _ada()
    sd = A.init_state_dict(res, 512, 512, nl, seed=0)
    G = gen.Generator(512, 512, nl, res, 3)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    ws0 = named_randn("ada:grad:w", (b, A.num_ws(res), 512), 5)
    gimg = named_randn("ada:grad:gy", (b, 3, res, res), 6)
    # oracle: fp64 autograd over ws and every parameter
    sd64 = {k: v.double().requires_grad_(v.is_floating_point() and "resampler" not in k) for k, v in sd.items()}
    ws64 = ws0.double().requires_grad_(True)
    ref = A.synthesis_network(sd64, res, ws64, "const")
    ref.backward(gimg.double())
    # frozen decoder, gradient w.r.t. the latents only
    for p in G.parameters():
        p.requires_grad_(False)
    ws = ws0.to(DEV).requires_grad_(True)
    img, _ = G([ws], input_is_latent=True, randomize_noise=False)
    with torch.no_grad():
        img_ng, _ = G([ws.detach()], input_is_latent=True, randomize_noise=False)
    assert (img - img_ng).abs().max() <= 1e-4 * img_ng.abs().max()              # both paths compute the same image
    assert (img.detach().cpu() - ref.detach().float()).abs().max() <= 2e-4 * ref.abs().max()
    img.backward(gimg.to(DEV))
    gref = ws64.grad.float()
    # fp32 vs fp64 through 7 leaky-relu layers: a pre-activation within rounding of the kink takes the other slope and
    # shifts every upstream gradient of that sample by ~1e-3 of the largest entry (measured; torch's own fp32 GPU
    # evaluation of the oracle, TF32 convolutions, is 2e-2 off).  Whole-network bounds are therefore above rounding;
    # the per-layer test below masks the kink and is tight.
    err = ws.grad.cpu() - gref
    assert err.abs().max() <= 1e-2 * gref.abs().max(), err.abs().max() / gref.abs().max()
    assert err.norm() <= 5e-3 * gref.norm(), err.norm() / gref.norm()
    # trainable decoder: every parameter gradient
    for n, p in G.named_parameters():
        p.requires_grad_("resampler" not in n)
    img, _ = G([ws0.to(DEV)], input_is_latent=True, randomize_noise=False)
    img.backward(gimg.to(DEV))
    checked = 0
    for n, p in G.named_parameters():
        if n.startswith("mapping") or "resampler" in n:
            continue
        r = sd64[n].grad.float()
        e = (p.grad.cpu() - r).norm().item()       # L2: one kink flip moves single entries of a weight gradient by percents
        tol = 3e-2 if p.numel() == 1 else 5e-3      # noise_strength: one fp32 sum of B*C*H*W cancelling terms
        assert e <= tol * r.norm().item() + 1e-6, (n, e, r.norm().item())
        checked += 1
    assert checked >= 4 + 13 * (len(A.block_resolutions(res)) - 1)


@pytest.mark.parametrize("cin,cout,res,up", [(16, 32, 8, False), (32, 16, 16, True), (8, 8, 4, False)])
def test_ada_layer_gradients_vs_oracle(sg2, cin, cout, res, up):
    """one SynthesisLayer2 / ToRGBLayer2, every gradient (x, w, parameters) vs fp64 autograd of the oracle; the
    upstream gradient is zeroed where the oracle's pre-activation is within 1e-3 of the leaky-relu kink, so the
    comparison is at rounding level."""
    from oracle import sg2_ada_oracle as A
    gen, U = _ada()
    g = torch.Generator().manual_seed(cin * 100 + res)
    b, rin = 3, res // 2 if up else res
    L = gen.SynthesisLayer2(cin, cout, 512, res, resampler=U.SmoothUpsample() if up else U.identity)
    T = gen.ToRGBLayer2(cin, 3, 512)
    with torch.no_grad():
        L.noise_strength.fill_(0.3)
        L.bias.copy_(0.2 * torch.randn(cout, generator=g))
        T.bias.copy_(0.2 * torch.randn(3, generator=g))
    x = torch.randn(b, cin, rin, rin, generator=g)
    w = torch.randn(b, 512, generator=g)
    for name, M in (("layer", L), ("torgb", T)):
        sd = {"L." + k: v.detach().double().requires_grad_("resampler" not in k and "noise_const" not in k)
              for k, v in M.state_dict().items()}
        x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
        if name == "layer":
            ref = A.synthesis_layer(sd, "L", x64, w64, sd["L.noise_const"], up)
            with torch.no_grad():                                     # the pre-activation, for the kink mask
                st = A.fully_connected(w64, sd["L.affine.weight"], sd["L.affine.bias"])
                pre = A.modulated_conv2d(x64, sd["L.weight"], st, padding=1)
                if up:
                    pre = A.smooth_upsample(pre, sd["L.resampler.kernel"])
                pre = pre + sd["L.noise_const"] * sd["L.noise_strength"] + sd["L.bias"][None, :, None, None]
            gy = torch.randn(ref.shape, generator=g) * (pre.abs() > 1e-3)
        else:
            ref = A.torgb_layer(sd, "L", x64, w64)
            gy = torch.randn(ref.shape, generator=g)
        ref.backward(gy.double())
        M = M.to(DEV)
        xd, wd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
        out = M(xd, wd, noise_mode="const") if name == "layer" else M(xd, wd)
        assert (out.detach().cpu() - ref.detach().float()).abs().max() <= 2e-5 * ref.abs().max()
        out.backward(gy.to(DEV))
        pairs = [("x", xd.grad, x64.grad), ("w", wd.grad, w64.grad)]
        pairs += [(k, p.grad, sd["L." + k].grad) for k, p in M.named_parameters() if "resampler" not in k]
        assert len(pairs) == (7 if name == "layer" else 6)
        for k, a, r in pairs:
            e = (a.cpu().double() - r).abs().max().item()
            assert e <= 1e-4 * r.abs().max().item() + 1e-7, (name, k, e, r.abs().max().item())


def test_ada_generator_bf16_precision(sg2):
    """precision='bf16': every 3x3 convolution on the tensor-core kernel (bf16 operands, fp32 accumulation), with and
    without autograd, vs the fp64 oracle"""
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    gen, U = _ada()
    res, b, nl = 32, 2, 2
    sd = A.init_state_dict(res, 512, 512, nl, seed=0)
    G = gen.Generator(512, 512, nl, res, 3)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    for p in G.parameters():
        p.requires_grad_(False)
    ws0 = named_randn("ada:grad:w", (b, A.num_ws(res), 512), 5)
    gimg = named_randn("ada:grad:gy", (b, 3, res, res), 6)
    ws64 = ws0.double().requires_grad_(True)
    ref = A.synthesis_network({k: v.double() for k, v in sd.items()}, res, ws64, "const")
    ref.backward(gimg.double())
    G.precision = 'bf16'
    assert G.synthesis.precision == 'bf16'
    n0 = sg2._lib.launch_count()
    with torch.no_grad():
        img, _ = G([ws0.to(DEV)], input_is_latent=True, randomize_noise=False)
    assert sg2._lib.launch_count() > n0
    assert (img.cpu().double() - ref.detach()).abs().max() <= 3e-2 * ref.abs().max()
    ws = ws0.to(DEV).requires_grad_(True)
    img2, _ = G([ws], input_is_latent=True, randomize_noise=False)
    assert (img2.detach() - img).abs().max() <= 3e-2 * ref.abs().max()
    img2.backward(gimg.to(DEV))
    err = ws.grad.cpu().double() - ws64.grad
    assert err.norm() <= 8e-2 * ws64.grad.norm(), (err.norm() / ws64.grad.norm()).item()
    G.precision = 'exact'
    with torch.no_grad():
        img3, _ = G([ws0.to(DEV)], input_is_latent=True, randomize_noise=False)
    assert (img3.cpu().double() - ref.detach()).abs().max() <= 2e-4 * ref.abs().max()


@pytest.mark.parametrize("res,b", [(16, 3), (64, 3), (256, 8), (1024, 2)])     # 1024: the 16-channel block runs zero-padded to 32
def test_ada_engine_vs_exact(sg2, res, b):
    """the ADA decoder on the whole-network bf16 engine (sg2_synth_create_ada: conv -> SmoothUpsample ordering, clamps,
    SmoothUpsample of the running image) against the exact fp32 path, itself pinned to the golden vectors of the
    unmodified reference module (test_ada_generator_golden) -- full images, per-sample worst case"""
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    gen, U = _ada()
    nl = 2 if res != 256 else 8
    sd = A.init_state_dict(res, 512, 512, nl, seed=0)
    G = gen.Generator(512, 512, nl, res, 3)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    z = named_randn(f"ada:eng:z{res}", (b, 512), 7).to(DEV)
    with torch.no_grad():
        G.precision = 'exact'
        ref, _ = G([z], randomize_noise=False)
        G.precision = 'bf16'
        assert G.synthesis._use_engine(torch.zeros(1, 4, 512, device=DEV), 'const')
        n0 = sg2._lib.launch_count()
        img, _ = G([z], randomize_noise=False)
        img2, _ = G([z], randomize_noise=False)
        assert "smoothup" in G.synthesis.engine().describe() and sg2._lib.launch_count() > n0
        rnd, _ = G([z], randomize_noise=True)
    assert img.shape == ref.shape and torch.equal(img, img2) and torch.isfinite(rnd).all() and not torch.equal(rnd, img)
    d = img.float() - ref.float()
    rel_max = (d.abs().amax(dim=(1, 2, 3)) / ref.abs().max()).max().item()
    rel_l2 = (d.flatten(1).norm(dim=1) / ref.float().flatten(1).norm(dim=1)).max().item()
    print(f"[parity] ADA engine {res}^2 B={b}: max|d|/|ref|max = {rel_max:.3e}, worst per-sample rel-L2 = {rel_l2:.3e}")
    assert rel_max <= 1.6e-2 and rel_l2 <= 2.4e-2, (rel_max, rel_l2)
