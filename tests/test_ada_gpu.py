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


def test_ada_generator_random_noise_and_grad_guard(sg2):
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
    with pytest.raises(RuntimeError, match="inference-only"):
        G([z.requires_grad_(True)], input_is_latent=False, randomize_noise=False)
