"""GPU: the training engine (csrc/synth_train.cu) -- the frozen decoder as one autograd node: forward = the launch plan with
kept activations, backward = the plan walked in reverse -- against fp64 autograd of the oracle (pinned to fp64 autograd of
the unmodified reference, tests/golden/grads.npz) and against the layer-by-layer autograd path of the same package."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(sg2, oracle, size, n_mlp):
    sd = oracle.init_state_dict(size, 512, n_mlp, 2, seed=0)
    G = sg2.Generator(size, 512, n_mlp)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    G.precision = "bf16"
    for p in G.parameters():
        p.requires_grad_(False)
    return G, sd


@pytest.mark.parametrize("size,B,per_sample_noise", [(16, 2, False), (32, 3, True), (64, 3, False), (256, 4, True)])
def test_train_engine_gradients_vs_oracle(sg2, oracle, size, B, per_sample_noise):
    n_mlp = 2 if size < 256 else 8
    G, sd = _gen(sg2, oracle, size, n_mlp)
    lat = 0.5 * oracle.named_randn(f"te:lat{size}", (B, G.n_latent, 512), 3)
    gy = oracle.named_randn(f"te:gy{size}", (B, 3, size, size), 3)
    noise = None
    if per_sample_noise:
        noise = [oracle.named_randn(f"te:n{size}:{i}", (B, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 3) for i in range(G.num_layers)]
    # oracle, fp64
    lo = lat.double().requires_grad_(True)
    kw = dict(noise=[n.double() for n in noise]) if noise else dict(randomize_noise=False)
    imgo, _ = oracle.generator_forward({k: v.double() for k, v in sd.items()}, size, [lo], n_mlp=n_mlp, input_is_latent=True, **kw)
    go, = torch.autograd.grad(imgo, [lo], gy.double())
    imgo = imgo.detach()
    # training engine
    kwd = dict(noise=[n.to(DEV) for n in noise]) if noise else dict(randomize_noise=False)
    ld = lat.to(DEV).requires_grad_(True)
    assert G._use_train_engine(ld.unsqueeze(0)[0], [None], False)
    img, _ = G([ld], input_is_latent=True, **kwd)
    assert type(img.grad_fn).__name__.startswith("SynthesisFunction")
    g_eng, = torch.autograd.grad(img, [ld], gy.to(DEV))
    with torch.no_grad():
        img_inf, _ = G([ld.detach()], input_is_latent=True, **kwd)
    assert torch.equal(img.detach(), img_inf)                       # the same launch plan produced both images
    e_img = ((img.detach().cpu().double() - imgo).abs().max() / imgo.abs().max()).item()
    e_g = ((g_eng.cpu().double() - go).norm() / go.norm()).item()
    # per latent row, so that one layer's style gradient cannot hide in the norm of the others
    row = ((g_eng.cpu().double() - go).flatten(2).norm(dim=2) / go.flatten(2).norm(dim=2).clamp_min(1e-30)).max().item()
    # layer-by-layer autograd path of the same package (tensor-core route), for scale
    os.environ["SG2_B200_TRAIN_ENGINE"] = "0"
    try:
        ld2 = lat.to(DEV).requires_grad_(True)
        img2, _ = G([ld2], input_is_latent=True, **kwd)
        g_old, = torch.autograd.grad(img2, [ld2], gy.to(DEV))
    finally:
        del os.environ["SG2_B200_TRAIN_ENGINE"]
    e_old = ((g_old.cpu().double() - go).norm() / go.norm()).item()
    print(f"[parity] train engine {size}^2 B={B}: image {e_img:.3e}, dL/dlatent rel-L2 {e_g:.3e} (worst latent row {row:.3e}); "
          f"layer-by-layer tensor-core path {e_old:.3e}")
    # bf16 operands forward and backward: the layer-by-layer tensor-core path measures 4-7e-2 on the same quantity
    assert e_img <= 1.6e-2 and e_g <= 1e-1 and row <= 2.5e-1, (e_img, e_g, row)


def test_train_engine_guards(sg2, oracle):
    G, sd = _gen(sg2, oracle, 16, 2)
    z = torch.randn(2, 512, device=DEV, requires_grad=True)        # through the mapping network and the latent broadcast
    img, _ = G([z], randomize_noise=False)
    img.sum().backward()
    assert z.grad is not None and torch.isfinite(z.grad).all() and z.grad.abs().max() > 0
    a = torch.randn(2, G.n_latent, 512, device=DEV, requires_grad=True)
    img1, _ = G([a], input_is_latent=True, randomize_noise=False)
    img2, _ = G([a], input_is_latent=True, randomize_noise=False)   # overwrites the kept activations of img1
    img2.sum().backward()
    with pytest.raises(RuntimeError, match="another forward"):
        img1.sum().backward()
    # anything the engine does not differentiate takes the layer-by-layer path
    n = [torch.randn(2, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device=DEV, requires_grad=True) for i in range(G.num_layers)]
    img3, _ = G([a], input_is_latent=True, noise=n)
    assert not type(img3.grad_fn).__name__.startswith("SynthesisFunction")
    G.conv1.conv.weight.requires_grad_(True)
    img4, _ = G([a], input_is_latent=True, randomize_noise=False)
    assert not type(img4.grad_fn).__name__.startswith("SynthesisFunction")


@pytest.mark.parametrize("size,cm,B", [(1024, 2, 1), (256, 1, 3), (512, 1, 2)])
def test_train_engine_vs_layerwise_route_large(sg2, oracle, size, cm, B):
    """the resolutions / widths whose forward uses the fused up-sampling conv, the merged polyphase walk and the dx-stacked
    kernel (1024^2; channel_multiplier 1): the engine's dL/dlatent against the layer-by-layer autograd route of the same
    package (itself checked against the oracle at 16^2..256^2) -- both carry bf16 operands, so they agree to a few percent"""
    sd = oracle.init_state_dict(size, 512, 2, cm, seed=0)
    G = sg2.Generator(size, 512, 2, channel_multiplier=cm)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    G.precision = "bf16"
    for p in G.parameters():
        p.requires_grad_(False)
    lat = 0.5 * oracle.named_randn(f"te:big:lat{size}", (B, G.n_latent, 512), 3).to(DEV)
    gy = oracle.named_randn(f"te:big:gy{size}", (B, 3, size, size), 3).to(DEV)
    ld = lat.clone().requires_grad_(True)
    img, _ = G([ld], input_is_latent=True, randomize_noise=False)
    assert type(img.grad_fn).__name__.startswith("SynthesisFunction")
    g_eng, = torch.autograd.grad(img, [ld], gy)
    os.environ["SG2_B200_TRAIN_ENGINE"] = "0"
    try:
        ld2 = lat.clone().requires_grad_(True)
        img2, _ = G([ld2], input_is_latent=True, randomize_noise=False)
        g_old, = torch.autograd.grad(img2, [ld2], gy)
    finally:
        del os.environ["SG2_B200_TRAIN_ENGINE"]
    e_img = ((img.detach() - img2.detach()).abs().max() / img2.detach().abs().max()).item()
    e_g = ((g_eng - g_old).norm() / g_old.norm()).item()
    row = ((g_eng - g_old).flatten(2).norm(dim=2) / g_old.flatten(2).norm(dim=2).clamp_min(1e-30)).max().item()
    print(f"[parity] train engine vs layer-by-layer route {size}^2 cm={cm} B={B}: image {e_img:.3e}, dL/dlatent rel-L2 {e_g:.3e} "
          f"(worst latent row {row:.3e})")
    assert torch.isfinite(g_eng).all() and e_img <= 2e-2 and e_g <= 1e-1 and row <= 2.5e-1, (e_img, e_g, row)
