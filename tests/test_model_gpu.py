"""GPU parity of the module path (exact fp32-accumulate kernels through the C ABI) against golden
vectors produced by the reference itself, and against the oracle on seeded inputs.

Tolerance (BASELINE.json north_star): fp32 outputs within 1e-3 max-abs of the reference's fp32
output.  The kernels reorder fp32 sums relative to cuDNN/oneDNN, so bit-exactness is not defined
for this floating-point path; measured differences are ~1e-5 (the reference's own fp32-vs-fp64
noise floor is 1.6e-5)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


def _layer(sg2, oracle, case):
    name, cin, cout, k, sdim, demod, up, down, (b, h, w) = case
    layer = sg2.ModulatedConv2d(cin, cout, k, sdim, demodulate=demod, upsample=up, downsample=down)
    layer.weight.data.copy_(oracle.named_randn(f"layer:{name}:weight", (1, cout, cin, k, k), 5))
    layer.modulation.weight.data.copy_(oracle.named_randn(f"layer:{name}:mod_w", (cin, sdim), 5))
    layer.modulation.bias.data.copy_(1 + 0.1 * oracle.named_randn(f"layer:{name}:mod_b", (cin,), 5))
    x = oracle.named_randn(f"layer:{name}:x", (b, cin, h, w), 5)
    s = oracle.named_randn(f"layer:{name}:s", (b, sdim), 5)
    return layer.to(DEV), x.to(DEV), s.to(DEV)


def test_modulated_conv_forward_golden(sg2, oracle, golden, cases):
    for case in cases.LAYER_CASES:
        layer, x, s = _layer(sg2, oracle, case)
        with torch.no_grad():
            y = layer(x, s)
        ref = golden["layers"][case[0] + "/y"]
        assert tuple(y.shape) == ref.shape, case[0]
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=0, atol=1e-4, err_msg=case[0])


def test_modulated_conv_backward_golden(sg2, oracle, golden, cases):
    for case in cases.LAYER_CASES:
        name = case[0]
        layer, x, s = _layer(sg2, oracle, case)
        x.requires_grad_(True), s.requires_grad_(True)
        y = layer(x, s)
        np.testing.assert_allclose(y.detach().cpu().numpy(), golden["layers"][name + "/y"], rtol=0, atol=1e-4)
        gy = oracle.named_randn(f"layer:{name}:gy", tuple(y.shape), 5).to(DEV)
        gx, gs, gw, gmw, gmb = torch.autograd.grad(
            y, [x, s, layer.weight, layer.modulation.weight, layer.modulation.bias], gy)
        for key, val in (("gx", gx), ("gs", gs), ("gw", gw), ("gmw", gmw), ("gmb", gmb)):
            ref = golden["layers"][f"{name}/{key}"]
            scale = max(1.0, float(np.abs(ref).max()))
            np.testing.assert_allclose(val.cpu().numpy(), ref, rtol=0, atol=2e-4 * scale, err_msg=f"{name}/{key}")


@pytest.mark.parametrize("cin,cout,res,up", [(512, 512, 4, False), (512, 512, 4, True), (512, 256, 16, True),
                                             (64, 64, 64, False), (32, 32, 128, False), (128, 3, 32, False)])
def test_modulated_conv_model_shapes_vs_oracle(sg2, oracle, cin, cout, res, up):
    k = 1 if cout == 3 else 3
    layer = sg2.ModulatedConv2d(cin, cout, k, 512, demodulate=cout != 3, upsample=up).to(DEV)
    x = torch.randn(2, cin, res, res, device=DEV)
    s = torch.randn(2, 512, device=DEV)
    with torch.no_grad():
        y = layer(x, s).cpu()
        taps = pad = None
        if up:
            taps, pad = oracle.fir_kernel_2d([1, 3, 3, 1]) * 4, oracle.upconv_blur_pad()
        ref = oracle.modulated_conv2d(x.cpu().double(), s.cpu().double(), layer.weight.cpu().double(),
                                      layer.modulation.weight.cpu().double(), layer.modulation.bias.cpu().double(),
                                      cout != 3, up, False, None if taps is None else taps.double(), pad or (0, 0))
    assert (y.double() - ref).abs().max() < 1e-4


def test_mapping_and_equal_linear_vs_oracle(sg2, oracle):
    sd = oracle.init_state_dict(16, 512, 8)
    G = sg2.Generator(16, 512, 8)
    G.load_state_dict(sd)
    G = G.to(DEV)
    for B in (1, 3, 64, 700):          # 700 > 4*148: the 8-samples-per-block variant
        z = oracle.named_randn(f"map:z{B}", (B, 512), 9)
        with torch.no_grad():
            w = G.style(z.to(DEV)).cpu()
            assert torch.equal(G.get_latent(z.to(DEV)).cpu(), w)
        ref = oracle.mapping_network({k: v.double() for k, v in sd.items()}, z.double(), 8).float()
        assert (w - ref).abs().max() <= 2e-4 * ref.abs().max()
    lin = sg2.EqualLinear(512, 96, bias_init=1, lr_mul=0.5).to(DEV)
    x = torch.randn(5, 7, 512, device=DEV)             # leading dims are batch
    with torch.no_grad():
        y = lin(x).cpu()
    ref = oracle.equal_linear(x.cpu().double(), lin.weight.detach().cpu().double(), lin.bias.detach().cpu().double(), 0.5)
    assert y.shape == (5, 7, 96) and (y.double() - ref).abs().max() < 1e-4
    m = G.mean_latent(4096)
    assert m.shape == (1, 512) and torch.isfinite(m).all()


def _run_case(sg2, oracle, cases, name, size, n_mlp, cm, batch, mode):
    sd = oracle.init_state_dict(size, 512, n_mlp, cm, seed=0)
    G = sg2.Generator(size, 512, n_mlp, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    G.precision = "exact"
    styles, kw = cases.gen_inputs(name, size, n_mlp, batch, mode, sd)
    styles = [s.to(DEV) for s in styles]
    kw = {k: ([n.to(DEV) for n in v] if k == "noise" else (v.to(DEV) if torch.is_tensor(v) else v)) for k, v in kw.items()}
    with torch.no_grad():
        img, aux = G(styles, **kw)
    return img.cpu(), None if aux is None else aux.cpu()


def test_generator_golden_small(sg2, oracle, golden, cases):
    for name, size, n_mlp, cm, batch, mode in cases.GEN_CASES:
        if size > 64:
            continue
        img, aux = _run_case(sg2, oracle, cases, name, size, n_mlp, cm, batch, mode)
        ref = torch.from_numpy(golden["generator"][name + "/img"])
        assert img.shape == ref.shape
        d = (img - ref).abs().max().item()
        assert d < TOL, f"{name}: max|d| {d:.2e} (|ref|max {ref.abs().max():.2f})"
        if aux is not None:
            refa = torch.from_numpy(golden["generator"][name + "/aux"])
            assert aux.shape == refa.shape and (aux - refa).abs().max() < TOL, name


def test_generator_golden_256(sg2, oracle, golden, cases):
    case = [c for c in cases.GEN_CASES if c[0] == "g256_z"][0]
    img, _ = _run_case(sg2, oracle, cases, *case)
    ref = torch.from_numpy(golden["generator"]["g256_z/img"])
    d = (img - ref).abs().max().item()
    assert img.shape == (1, 3, 256, 256) and d < TOL, f"max|d| {d:.2e}, |ref|max {ref.abs().max():.2f}"


def test_generator_golden_1024(sg2, oracle, golden, cases):
    case = [c for c in cases.GEN_CASES if c[0] == "g1024_wplus"][0]
    img, latent = _run_case(sg2, oracle, cases, *case)
    assert img.shape == (1, 3, 1024, 1024) and latent.shape == (1, 18, 512)
    ref = torch.from_numpy(golden["generator"]["g1024_wplus/img_lattice8"])
    d = (img[:, :, 3::8, 5::8] - ref).abs().max().item()
    assert d < TOL, f"max|d| {d:.2e}"
    mom = golden["generator"]["g1024_wplus/img_moments"]
    assert abs(img.double().mean().item() - mom[0]) < 1e-4 and abs(img.double().std().item() - mom[1]) < 1e-4
    assert abs(img.min().item() - mom[3]) < TOL and abs(img.max().item() - mom[4]) < TOL


def test_generator_api_surface(sg2, oracle):
    sd = oracle.init_state_dict(16, 512, 2)
    G = sg2.Generator(16, 512, 2)
    G.load_state_dict(sd)
    G = G.to(DEV).eval()
    z = torch.randn(3, 512, device=DEV)
    with torch.no_grad():
        out = G([z])                                         # randomize_noise=True default
        assert isinstance(out, tuple) and len(out) == 2 and out[1] is None and out[0].shape == (3, 3, 16, 16)
        a, _ = G([z], randomize_noise=False)
        b, _ = G([z], randomize_noise=False)
        assert torch.equal(a, b)                             # deterministic without random noise
        c, _ = G([z], randomize_noise=True)
        assert not torch.equal(a, c)                         # fresh noise is drawn (noise weights are non-zero)
        img, lat = G([z], return_latents=True, randomize_noise=False)
        assert lat.shape == (3, G.n_latent, 512)
        img2, _ = G([lat], input_is_latent=True, randomize_noise=False)
        assert torch.equal(img, img2)                        # w+ passthrough (psp.py:108-111 call)
        _, feat = G([z], return_features=True, randomize_noise=False)
        assert feat.shape == (3, 512, 16, 16)
        mix, lat2 = G([z, z.flip(0)], return_latents=True, randomize_noise=False)   # random inject_index
        assert lat2.shape == lat.shape
        noise = G.make_noise()
        n1, _ = G([z], noise=noise)
        n2, _ = G([z], noise=noise)
        assert torch.equal(n1, n2)
    # half precision storage runs through the same kernels (fp32 accumulate)
    Gh = sg2.Generator(16, 512, 2)
    Gh.load_state_dict(sd)
    Gh = Gh.to(DEV).half().eval()
    Gh.precision = "exact"
    with torch.no_grad():
        h, _ = Gh([z.half()], randomize_noise=False)
    assert h.dtype == torch.float16 and (h.float() - a).abs().max() < 0.05 * a.abs().max()


def test_generator_gradients_vs_oracle(sg2, oracle):
    """dL/dlatent and dL/dnoise through the whole decoder (the ReStyle fine-tuning direction)."""
    sd = oracle.init_state_dict(16, 512, 2)
    G = sg2.Generator(16, 512, 2)
    G.load_state_dict(sd)
    G = G.to(DEV).eval()
    lat = (0.5 * oracle.named_randn("grad:lat", (2, 6, 512), 3))
    noise = [oracle.named_randn(f"grad:n{i}", (2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 3) for i in range(5)]
    gy = oracle.named_randn("grad:gy", (2, 3, 16, 16), 3)
    ld = lat.to(DEV).requires_grad_(True)
    nd = [n.to(DEV).requires_grad_(True) for n in noise]
    img, _ = G([ld], input_is_latent=True, noise=nd)
    grads = torch.autograd.grad(img, [ld] + nd, gy.to(DEV))
    lo = lat.double().requires_grad_(True)
    no = [n.double().requires_grad_(True) for n in noise]
    imgo, _ = oracle.generator_forward({k: v.double() for k, v in sd.items()}, 16, [lo], n_mlp=2,
                                       input_is_latent=True, noise=no)
    assert (img.detach().cpu().double() - imgo.detach()).abs().max() < TOL
    go = torch.autograd.grad(imgo, [lo] + no, gy.double())
    for a, b in zip(grads, go):
        scale = max(1.0, b.abs().max().item())
        assert (a.cpu().double() - b).abs().max() < 1e-3 * scale


def test_discriminator_matches_reference_golden(sg2):
    """the Discriminator mirror on the kernels vs golden vectors of the unmodified reference class
    (tests/golden/make_golden_disc.py); the differentiable route, library convolutions in true fp32"""
    import os
    import numpy as np
    import make_golden_disc as MD
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "disc.npz"))
    for name, size, cm, batch in MD.CASES:
        D = sg2.Discriminator(size, channel_multiplier=cm).eval()
        D.load_state_dict(MD.seeded_state_dict(D, name), strict=True)
        D = D.to(DEV)
        x = MD.images(name, batch, size).to(DEV).requires_grad_(True)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            y = D(x)
        ref = torch.from_numpy(g[name + "/out"])
        err = (y.detach().cpu() - ref).abs().max().item()
        assert y.shape == ref.shape and err <= 1e-3 * max(1.0, ref.abs().max().item()), (name, err)


def test_discriminator_runs_on_the_op_api(sg2):
    D = sg2.Discriminator(32).to(DEV)
    x = torch.randn(4, 3, 32, 32, device=DEV, requires_grad=True)
    out = D(x)
    assert out.shape == (4, 1)
    out.sum().backward()
    assert torch.isfinite(x.grad).all()
