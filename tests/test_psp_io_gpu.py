"""GPU: face_pool / bilinear resize after the decoder vs the PyTorch ops the reference calls (psp.py:33,
coach_restyle_psp.py:156), evaluated on CPU."""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_face_pool_and_resize_vs_torch(sg2):
    io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
    g = torch.Generator().manual_seed(0)
    for s in (256, 512, 1024):
        x = torch.randn(2, 3, s, s, generator=g)
        ref = torch.nn.AdaptiveAvgPool2d((256, 256))(x)
        y = io.FacePool((256, 256))(x.to(DEV)).cpu()
        np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=2e-6)
    for shape, size in (((2, 3, 256, 256), 112), ((1, 2, 37, 53), (20, 71)), ((1, 1, 4, 4), 9), ((2, 3, 192, 256), 112)):
        x = torch.randn(shape, generator=g)
        ref = F.interpolate(x, size, mode="bilinear")
        y = io.resize_bilinear(x.to(DEV), size).cpu()
        assert y.shape == ref.shape
        np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=3e-6)
    x = torch.randn(2, 3, 1024, 1024, generator=g)
    pooled, small = io.decode_epilogue(x.to(DEV))
    ref = F.interpolate(torch.nn.AdaptiveAvgPool2d((256, 256))(x), 112, mode="bilinear")
    np.testing.assert_allclose(small.cpu().numpy(), ref.numpy(), rtol=0, atol=3e-6)
    xb = x.to(DEV).bfloat16()
    sb = io.resize_bilinear(io.face_pool(xb), 112).float().cpu()
    assert (sb - ref).abs().max() < 3e-2
    with pytest.raises(RuntimeError):
        io.face_pool(torch.randn(1, 3, 300, 300, device=DEV))
    with pytest.raises(RuntimeError, match="CUDA"):
        io.face_pool(torch.randn(1, 3, 512, 512))


def test_images_to_uint8(sg2):
    io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
    x = torch.randn(3, 3, 37, 41) * 0.8
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -3.0, 5.0])
    ref = ((x.numpy() + 1) / 2).clip(0, 1) * 255
    y = io.images_to_uint8(x.to(DEV)).cpu().numpy()
    assert y.dtype == np.uint8 and y.shape == x.shape
    assert np.abs(y.astype(np.int32) - ref.astype("uint8").astype(np.int32)).max() <= 1     # fp32 rounding at bin edges
    assert (y == ref.astype("uint8")).mean() > 0.999


def test_face_pool_and_resize_gradients_vs_torch(sg2):
    """the coaches back-propagate the image losses through both ops (coach_restyle_psp.py:86-101,143-156)"""
    io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
    g = torch.Generator().manual_seed(1)
    for s in (512, 1024):
        x = torch.randn(1, 3, s, s, generator=g)
        gy = torch.randn(1, 3, 256, 256, generator=g)
        xr = x.clone().requires_grad_(True)
        torch.nn.AdaptiveAvgPool2d((256, 256))(xr).backward(gy)
        xd = x.to(DEV).requires_grad_(True)
        io.face_pool(xd).backward(gy.to(DEV))
        np.testing.assert_allclose(xd.grad.cpu().numpy(), xr.grad.numpy(), rtol=0, atol=1e-7)
    for shape, size in (((2, 3, 256, 256), 112), ((1, 2, 37, 53), (20, 71)), ((1, 1, 4, 4), 9), ((1, 1, 9, 9), 4),
                        ((2, 3, 192, 256), 112), ((1, 1, 1, 5), (3, 5)), ((1, 2, 16, 16), 64), ((1, 1, 7, 7), 7)):
        x = torch.randn(shape, generator=g)
        oh, ow = (size, size) if isinstance(size, int) else size
        gy = torch.randn(shape[0], shape[1], oh, ow, generator=g)
        xr = x.clone().requires_grad_(True)
        F.interpolate(xr, size, mode="bilinear").backward(gy)
        xd = x.to(DEV).requires_grad_(True)
        y = io.resize_bilinear(xd, size)
        y.backward(gy.to(DEV))
        np.testing.assert_allclose(xd.grad.cpu().numpy(), xr.grad.numpy(), rtol=0, atol=5e-6, err_msg=str((shape, size)))
    # chained, low precision storage
    x = torch.randn(1, 3, 1024, 1024, generator=g)
    gy = torch.randn(1, 3, 112, 112, generator=g)
    xr = x.clone().requires_grad_(True)
    F.interpolate(torch.nn.AdaptiveAvgPool2d((256, 256))(xr), 112, mode="bilinear").backward(gy)
    xd = x.to(DEV).bfloat16().requires_grad_(True)
    pooled, small = io.decode_epilogue(xd)
    small.backward(gy.to(DEV).bfloat16())
    assert xd.grad.dtype == torch.bfloat16
    assert (xd.grad.float().cpu() - xr.grad).abs().max() <= 2e-2 * xr.grad.abs().max()
    with pytest.raises(RuntimeError, match="inference-only"):
        io.images_to_uint8(xd)


@pytest.mark.parametrize("size,pooled,b", [(64, 16, 3), (64, 32, 3), (1024, 256, 2), (512, 256, 2)])
def test_decode_pooled_matches_decoder_then_face_pool(sg2, size, pooled, b):
    """psp.py:104-114 as one call: the engine's last launch writes the face-pooled image itself (no full-resolution image is
    stored) -- equal to the same engine's image pooled by the op kernel, through the CUDA graph and without it, from z and
    from w+; on the exact path (no engine) the two steps simply run one after the other"""
    io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
    torch.manual_seed(0)
    G = sg2.Generator(size, 512, 2).to(DEV).eval()
    z = torch.randn(b, 512, device=DEV)
    with torch.no_grad():
        G.precision = "bf16"
        full, lat = G([z], randomize_noise=False, return_latents=True)
        ref = io.face_pool(full, pooled)
        n0 = sg2._lib.launch_count()
        y, lat2 = io.decode_pooled(G, [z], pooled, randomize_noise=False, return_latents=True)          # from z, graph
        assert sg2._lib.launch_count() > n0
        assert y.shape == (b, 3, pooled, pooled) and torch.equal(lat, lat2)
        scale = float(ref.abs().max())
        assert float((y - ref).abs().max()) <= 2e-6 * max(scale, 1.0)
        y2, _ = io.decode_pooled(G, [lat], pooled, input_is_latent=True, randomize_noise=False)         # from w+, graph
        assert float((y2 - ref).abs().max()) <= 2e-6 * max(scale, 1.0)
        eng = G.engine()
        eng.use_graph = False
        y3, _ = io.decode_pooled(G, [lat], pooled, input_is_latent=True, randomize_noise=False)         # plain launches
        eng.use_graph = True
        assert float((y3 - ref).abs().max()) <= 2e-6 * max(scale, 1.0)
        full_again, _ = G([z], randomize_noise=False)                # the request does not leak into ordinary forwards
        assert full_again.shape[-1] == size and torch.equal(full_again, full)
        if size <= 64:
            G.precision = "exact"
            fe, _ = G([z], randomize_noise=False)
            ye, _ = io.decode_pooled(G, [z], pooled, randomize_noise=False)
            assert torch.equal(ye, io.face_pool(fe, pooled))
