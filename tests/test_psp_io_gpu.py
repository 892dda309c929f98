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
