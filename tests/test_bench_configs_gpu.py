"""GPU parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs[1], configs[3] and the fine-tuning direction of
configs[2]): the shapes bench.py times are the shapes compared here, full images, not a lattice.

* bf16 tcgen05 engine vs the exact fp32 path (itself pinned to the reference's golden vectors at the same resolutions,
  tests/test_model_gpu.py::test_generator_golden_{256,1024}) at 256^2 B = 64 and 1024^2 B = 32, with the registered
  (broadcast) noise buffers and with one noise map per sample -- a multi-sample-block tile walk at full resolution.
* whole-decoder gradients w.r.t. the latents and every noise map at 256^2, B = 4, both decoders, vs fp64 autograd of the
  oracle (pinned to fp64 autograd of the unmodified reference, tests/golden/grads.npz).

Tolerances are <= 2x the values measured on B200 (printed by every test; profiles/pytest_gpu_r02_*.log)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# measured on B200 (round 2, profiles/pytest_gpu_r02_*.log): max|d|/|ref|max 8.3e-3 and worst per-sample rel-L2 1.08e-2 at
# 256^2 B = 64; 8.4e-3 and 1.24e-2 at 1024^2 B = 32 (bf16 activations and weights through 13 / 17 layers)
ENGINE_REL_MAX, ENGINE_REL_L2 = 1.6e-2, 2.4e-2


def _gen(sg2, oracle, size):
    sd = oracle.init_state_dict(size, 512, 8, 2, seed=0)
    G = sg2.Generator(size, 512, 8)
    G.load_state_dict(sd, strict=True)
    return G.to(DEV).eval(), sd


def _compare(img, ref, what):
    assert img.shape == ref.shape and torch.isfinite(img).all(), what
    scale = ref.abs().max()
    # per-sample worst case, so that one bad sample block cannot hide in the batch norm
    d = (img.float() - ref.float())
    rel_max = (d.abs().amax(dim=(1, 2, 3)) / scale).max().item()
    rel_l2 = (d.flatten(1).norm(dim=1) / ref.float().flatten(1).norm(dim=1)).max().item()
    print(f"[parity] {what}: max|d|/|ref|max = {rel_max:.3e}, worst per-sample rel-L2 = {rel_l2:.3e}")
    assert rel_max <= ENGINE_REL_MAX and rel_l2 <= ENGINE_REL_L2, (what, rel_max, rel_l2)


@pytest.mark.parametrize("size,B", [(256, 64), (1024, 32)])
def test_engine_vs_exact_at_bench_config(sg2, oracle, size, B):
    G, _ = _gen(sg2, oracle, size)
    z = oracle.named_randn(f"benchcfg:z{size}", (B, 512), 11).to(DEV)
    noise = [oracle.named_randn(f"benchcfg:n{size}:{i}", (B, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 11).to(DEV)
             for i in range(G.num_layers)]
    with torch.no_grad():
        for what, kw in (("registered noise buffers", dict(randomize_noise=False)), ("one noise map per sample", dict(noise=noise))):
            G.precision = "exact"
            ref, _ = G([z], **kw)
            G.precision = "bf16"
            img, _ = G([z], **kw)
            _compare(img, ref, f"{size}^2 B={B}, {what}")
            img2, _ = G([z], **kw)
            assert torch.equal(img, img2), "engine is not deterministic"
            del ref, img, img2
    # w+ latents (the pSp decode call, psp.py:108) through the same plan
    with torch.no_grad():
        w = G.style(z)
        wp = w.unsqueeze(1).repeat(1, G.n_latent, 1) + 0.1 * oracle.named_randn(f"benchcfg:wp{size}", (B, G.n_latent, 512), 11).to(DEV)
        G.precision = "exact"
        ref, _ = G([wp], input_is_latent=True, randomize_noise=False)
        G.precision = "bf16"
        img, _ = G([wp], input_is_latent=True, randomize_noise=False)
        _compare(img, ref, f"{size}^2 B={B}, w+ latents")


def test_rosinality_decoder_gradients_256(sg2, oracle):
    """dL/dlatent and dL/dnoise through the 256^2 decoder at B = 4 (fine-tuning direction, coach_restyle_psp.py:138-168):
    exact fp32 route and tensor-core route vs fp64 autograd of the oracle."""
    size, B = 256, 4
    G, sd = _gen(sg2, oracle, size)
    for p in G.parameters():
        p.requires_grad_(False)
    lat = 0.5 * oracle.named_randn("benchcfg:grad:lat", (B, G.n_latent, 512), 3)
    noise = [oracle.named_randn(f"benchcfg:grad:n{i}", (B, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 3) for i in range(G.num_layers)]
    gy = oracle.named_randn("benchcfg:grad:gy", (B, 3, size, size), 3)
    lo = lat.double().requires_grad_(True)
    no = [n.double().requires_grad_(True) for n in noise]
    imgo, _ = oracle.generator_forward({k: v.double() for k, v in sd.items()}, size, [lo], n_mlp=8, input_is_latent=True, noise=no)
    go = torch.autograd.grad(imgo, [lo] + no, gy.double())
    imgo = imgo.detach()
    # measured on B200 (round 2): exact route image 2.5e-6, dL/dlatent rel-L2 2.0e-3, worst dL/dnoise 4.2e-3 (fp32 vs fp64:
    # pre-activations within rounding of the leaky-relu kink take the other slope, see test_ada_gpu.py)
    for precision, tol_img, tol_lat, tol_noise in (("exact", 1e-3, 4e-3, 8.4e-3), ("bf16", 1.6e-2, 1e-1, 1e-1)):
        G.precision = precision
        ld = lat.to(DEV).requires_grad_(True)
        nd = [n.to(DEV).requires_grad_(True) for n in noise]
        img, _ = G([ld], input_is_latent=True, noise=nd)
        grads = torch.autograd.grad(img, [ld] + nd, gy.to(DEV))
        e_img = ((img.detach().cpu().double() - imgo).abs().max() / imgo.abs().max()).item()
        e_lat = ((grads[0].cpu().double() - go[0]).norm() / go[0].norm()).item()
        e_noise = max(((a.cpu().double() - b).norm() / b.norm()).item() for a, b in zip(grads[1:], go[1:]))
        print(f"[parity] rosinality 256^2 B=4 gradients, {precision}: image {e_img:.3e}, dL/dlatent rel-L2 {e_lat:.3e}, "
              f"worst dL/dnoise rel-L2 {e_noise:.3e}")
        assert e_img <= tol_img and e_lat <= tol_lat and e_noise <= tol_noise, (precision, e_img, e_lat, e_noise)


def test_ada_decoder_gradients_256(sg2):
    """the ADA decoder (the one psp.py:24-30 builds for the documented Stage-2 command) at 256^2, B = 4."""
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    gen = importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator")
    res, B, nl = 256, 4, 8
    sd = A.init_state_dict(res, 512, 512, nl, seed=0)
    G = gen.Generator(512, 512, nl, res, 3)
    G.load_state_dict(sd, strict=True)
    G = G.to(DEV).eval()
    for p in G.parameters():
        p.requires_grad_(False)
    ws0 = named_randn("benchcfg:ada:w", (B, A.num_ws(res), 512), 5)
    gimg = named_randn("benchcfg:ada:gy", (B, 3, res, res), 6)
    ws64 = ws0.double().requires_grad_(True)
    ref = A.synthesis_network({k: v.double() for k, v in sd.items()}, res, ws64, "const")
    ref.backward(gimg.double())
    gref = ws64.grad
    ref = ref.detach()
    # measured on B200 (round 2): exact image 2.3e-6, dL/dw rel-L2 6.0e-4 (leaky-relu kinks, see test_ada_gpu.py); bf16 7.4e-3 / 4.7e-2
    for precision, tol_img, tol_g in (("exact", 2e-4, 1.2e-3), ("bf16", 1.5e-2, 9.5e-2)):
        G.precision = precision
        ws = ws0.to(DEV).requires_grad_(True)
        img, _ = G([ws], input_is_latent=True, randomize_noise=False)
        img.backward(gimg.to(DEV))
        e_img = ((img.detach().cpu().double() - ref).abs().max() / ref.abs().max()).item()
        e_g = ((ws.grad.cpu().double() - gref).norm() / gref.norm()).item()
        print(f"[parity] ADA 256^2 B=4 gradients, {precision}: image {e_img:.3e}, dL/dw rel-L2 {e_g:.3e}")
        assert e_img <= tol_img and e_g <= tol_g, (precision, e_img, e_g)
