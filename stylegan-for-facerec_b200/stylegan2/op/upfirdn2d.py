"""upfirdn2d over the sm_100a kernel `sg2_upfirdn2d`.

Drop-in for the reference's op/upfirdn2d.py (UpFirDn2dBackward :17-84, UpFirDn2d :87-139,
upfirdn2d :142-147): same call signature, output size, autograd (first and second order; the FIR
taps get no gradient, as in the reference).  Differences, all supersets: any up/down <= 4 and taps
<= 16x16 instead of six hard-coded modes, an error instead of uninitialised memory outside that,
fp32 accumulation for half/bfloat16 storage, bfloat16 support.
"""
import torch
from torch.autograd import Function

from ... import _lib


def out_size(n, up, down, p0, p1, k):
    return (n * up + p0 + p1 - k) // down + 1          # op/upfirdn2d.py:100-101


def upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """x4 [major, in_h, in_w, minor] -> [major, out_h, out_w, minor]; the 10-argument contract of the
    reference's pybind `upfirdn2d.upfirdn2d` (op/upfirdn2d.cpp:12-23)."""
    _lib.require_cuda(x4)
    lib = _lib.load()
    x4 = x4.contiguous()
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    taps = kernel.detach().to(device=x4.device, dtype=torch.float32).contiguous()
    oh, ow = out_size(in_h, up_y, down_y, py0, py1, kh), out_size(in_w, up_x, down_x, px0, px1, kw)
    if oh < 1 or ow < 1:
        raise RuntimeError(f"upfirdn2d: empty output ({oh} x {ow})")
    out = torch.empty((major, oh, ow, minor), device=x4.device, dtype=x4.dtype)
    with _lib.device_of(x4):
        _lib.check(lib.sg2_upfirdn2d(out.data_ptr(), x4.data_ptr(), taps.data_ptr(), major, in_h, in_w, minor,
                                     kh, kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1,
                                     _lib.dtype_code(x4), _lib.stream_of(x4)), "upfirdn2d")
    return out


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size_):
        up_x, up_y = up
        down_x, down_y = down
        gx0, gx1, gy0, gy1 = g_pad
        g = grad_output.reshape(-1, out_size_[0], out_size_[1], 1)
        # adjoint = same op with up/down swapped and flipped taps (op/upfirdn2d.py:28-41)
        gi = upfirdn2d_raw(g, grad_kernel, down_x, down_y, up_x, up_y, gx0, gx1, gy0, gy1)
        gi = gi.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.cfg = (up_x, up_y, down_x, down_y) + tuple(pad)
        ctx.in_size, ctx.out_size = in_size, out_size_
        return gi

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gg = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1)
        ggo = upfirdn2d_raw(gg, kernel, *ctx.cfg)
        ggo = ggo.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return ggo, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        _, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        x4 = input.reshape(-1, in_h, in_w, 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        oh, ow = out_size(in_h, up_y, down_y, py0, py1, kh), out_size(in_w, up_x, down_x, px0, px1, kw)
        ctx.out_size = (oh, ow)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (px0, px1, py0, py1)
        # pads of the adjoint (op/upfirdn2d.py:108-113)
        ctx.g_pad = (kw - px0 - 1, in_w * up_x - ow * down_x + px0 - up_x + 1,
                     kh - py0 - 1, in_h * up_y - oh * down_y + py0 - up_y + 1)
        out = upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
        return out.view(-1, channel, oh, ow)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        gi = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                     ctx.g_pad, ctx.in_size, ctx.out_size)
        return gi, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
