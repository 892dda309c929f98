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
