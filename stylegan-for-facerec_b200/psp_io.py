"""The step right after the decoder in ReStyle / pSp on the sg2_b200 kernels:

    images = net.face_pool(images)                       # psp.py:33,113-114  AdaptiveAvgPool2d((256, 256))
    y_hat = F.interpolate(y_hat, 112, mode='bilinear')   # coach_restyle_psp.py:156

`FacePool` is a drop-in for the `torch.nn.AdaptiveAvgPool2d((256, 256))` the reference assigns to
`pSp.face_pool`; `resize_bilinear` for the interpolate call.  Both carry autograd (forward kernel + its exact
adjoint kernel, csrc/psp_ops.cu) because the coaches back-propagate the image losses through them;
`images_to_uint8` is an output conversion and has none."""
import torch

from . import _lib
from .stylegan2.functional import needs_grad


def _guard(x, what):
    _lib.require_cuda(x)
    if needs_grad(x):
        raise RuntimeError(f"sg2_b200 {what}: inference-only kernel (no autograd yet); detach the input or use torch.no_grad()")


class _FacePoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow, f):
        ctx.geom = (oh, ow, f)
        return _face_pool_fwd(x.detach(), oh, ow, f)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        oh, ow, f = ctx.geom
        gy = gy.contiguous()
        b, c = gy.shape[:2]
        gx = torch.empty((b, c, oh * f, ow * f), device=gy.device, dtype=gy.dtype)
        with _lib.device_of(gy):
            _lib.check(_lib.load().sg2_avg_pool_int_bwd(gx.data_ptr(), gy.data_ptr(), b * c, oh, ow, f, _lib.dtype_code(gy),
                                                        _lib.stream_of(gy)), "avg_pool_int_bwd")
        return gx, None, None, None


class _ResizeBilinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow):
        ctx.geom = (x.shape[2], x.shape[3], oh, ow)
        return _resize_fwd(x.detach(), oh, ow)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        h, w, oh, ow = ctx.geom
        gy = gy.contiguous()
        b, c = gy.shape[:2]
        gx = torch.empty((b, c, h, w), device=gy.device, dtype=gy.dtype)
        with _lib.device_of(gy):
            _lib.check(_lib.load().sg2_resize_bilinear_bwd(gx.data_ptr(), gy.data_ptr(), b * c, h, w, oh, ow,
                                                           _lib.dtype_code(gy), _lib.stream_of(gy)), "resize_bilinear_bwd")
        return gx, None, None


def _face_pool_fwd(images, oh, ow, f):
    b, c = images.shape[:2]
    x = images.contiguous()
    out = torch.empty((b, c, oh, ow), device=x.device, dtype=x.dtype)
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_avg_pool_int(out.data_ptr(), x.data_ptr(), b * c, oh, ow, f, _lib.dtype_code(x),
                                                _lib.stream_of(x)), "avg_pool_int")
    return out


def _resize_fwd(images, oh, ow):
    b, c, h, w = images.shape
    x = images.contiguous()
    out = torch.empty((b, c, oh, ow), device=x.device, dtype=x.dtype)
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_resize_bilinear(out.data_ptr(), x.data_ptr(), b * c, h, w, oh, ow, _lib.dtype_code(x),
                                                   _lib.stream_of(x)), "resize_bilinear")
    return out


def face_pool(images: torch.Tensor, size=(256, 256)) -> torch.Tensor:
    """AdaptiveAvgPool2d(size) for integer pooling ratios (1024 -> 256, 512 -> 256, 256 -> 256)."""
    _lib.require_cuda(images)
    oh, ow = (size, size) if isinstance(size, int) else size
    b, c, h, w = images.shape
    if h % oh or w % ow or h // oh != w // ow:
        raise RuntimeError(f"sg2_b200 face_pool: {h}x{w} -> {oh}x{ow} is not an integer, isotropic ratio")
    f = h // oh
    if f == 1:
        return images
    if needs_grad(images):
        return _FacePoolFunction.apply(images, oh, ow, f)
    return _face_pool_fwd(images, oh, ow, f)


def resize_bilinear(images: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(images, size, mode='bilinear') with the PyTorch defaults (align_corners=False, no antialias)."""
    _lib.require_cuda(images)
    oh, ow = (size, size) if isinstance(size, int) else size
    if needs_grad(images):
        return _ResizeBilinearFunction.apply(images, oh, ow)
    return _resize_fwd(images, oh, ow)


class FacePool(torch.nn.Module):
    """drop-in for `self.face_pool = torch.nn.AdaptiveAvgPool2d((256, 256))` (psp.py:33)"""

    def __init__(self, output_size=(256, 256)):
        super().__init__()
        self.output_size = output_size

    def forward(self, x):
        return face_pool(x, self.output_size)


def decode_pooled(generator, styles, size=(256, 256), **forward_kwargs):
    """`images, latent = decoder(styles, **kw); images = face_pool(images)` (psp.py:104-114) as one call: when the decoder runs
    on the whole-network engine (bf16, nothing requires grad) and the pooling ratio is 2 or 4, the LAST launch of the forward
    writes the pooled image itself and the full-resolution image is never stored; otherwise the two steps run one
    after the other (`face_pool` carries autograd).  -> (pooled images, latent or None)"""
    oh, ow = (size, size) if isinstance(size, int) else size
    full = getattr(generator, "size", None)
    factor = full // oh if (full and oh == ow and full % oh == 0) else 0
    if factor in (2, 4):
        generator.__dict__["_pool_request"] = factor
    try:
        images, latent = generator(styles, **forward_kwargs)
    finally:
        generator.__dict__.pop("_pool_request", None)
    if images.shape[-2:] != (oh, ow):            # the engine did not take the request (autograd, exact path, other ratios)
        images = face_pool(images, (oh, ow))
    return images, latent


def decode_epilogue(images: torch.Tensor, pool=(256, 256), resize=112):
    """decoder output -> (face-pooled image for the next refinement step, its 112x112 version for the losses)."""
    pooled = face_pool(images, pool)
    return pooled, resize_bilinear(pooled, resize)


def images_to_uint8(images: torch.Tensor) -> torch.Tensor:
    """tensor2im on the device (utils/common.py:5-11 of the reference: ((x + 1) / 2).clip(0, 1) * 255 -> uint8),
    same NCHW layout: a quarter of the bytes to bring back to the host / write to disk."""
    _guard(images, "images_to_uint8")
    x = images.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_image_to_uint8(out.data_ptr(), x.data_ptr(), x.numel(), _lib.dtype_code(x),
                                                  _lib.stream_of(x)), "image_to_uint8")
    return out
