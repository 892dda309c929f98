"""Host-side wrappers of the C-ABI kernels used by model.py (exact NCHW path) and the shared-weight convolution
`Function`; the tensor-core route of the differentiable path lives in tc_route.py and is re-exported here.

Each wrapper makes inputs contiguous, allocates outputs/workspace with torch's caching allocator
and passes raw pointers + the current stream; the C side owns nothing.
"""
import math

import torch
import ctypes as C

from .. import _lib


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def equal_linear(x, weight, bias, w_scale, lr_mul, act):
    """model.py:147-157 without autograd. x [..., in] -> [..., out]."""
    _lib.require_cuda(x)
    lib = _lib.load()
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    w = weight.detach().to(x2.dtype).contiguous()
    b = None if bias is None else bias.detach().to(x2.dtype).contiguous()
    out = torch.empty((x2.shape[0], w.shape[0]), device=x.device, dtype=x2.dtype)
    with _lib.device_of(x2):
        _lib.check(lib.sg2_equal_linear_fwd(out.data_ptr(), x2.data_ptr(), w.data_ptr(), _lib.ptr(b),
                                            x2.shape[0], w.shape[1], w.shape[0], float(w_scale), float(lr_mul),
                                            1 if act else 0, _lib.dtype_code(x2), _lib.stream_of(x2)),
                   "equal_linear_fwd")
    return out.reshape(*lead, w.shape[0])


def mapping(z, weights, biases, lr_mul, pixel_norm=True):
    """Generator.style in one launch (model.py:10-15,378-387). z [B, dim]."""
    _lib.require_cuda(z)
    lib = _lib.load()
    z2 = z.contiguous()
    n = len(weights)
    ws = [w.detach().to(z2.dtype).contiguous() for w in weights]
    bs = [b.detach().to(z2.dtype).contiguous() for b in biases]
    wp = (C.c_void_p * max(n, 1))(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * max(n, 1))(*[b.data_ptr() for b in bs])
    out = torch.empty_like(z2)
    scratch = torch.empty_like(z2) if n > 1 else None
    with _lib.device_of(z2):
        _lib.check(lib.sg2_mapping_fwd(out.data_ptr(), z2.data_ptr(), wp, bp, n, z2.shape[0], z2.shape[1],
                                       float(lr_mul), 1 if pixel_norm else 0, _lib.ptr(scratch),
                                       _lib.dtype_code(z2), _lib.stream_of(z2)), "mapping_fwd")
    return out


def conv_prep(weight4, conv_scale, want_wsq=True):
    """weight4 [Cout,Cin,k,k] -> (wt fp32 [Cin*k*k, Cout] scaled, wsq fp32 [Cin, Cout] or None)."""
    lib = _lib.load()
    w = weight4.detach().contiguous()
    cout, cin, k, _ = w.shape
    wt = torch.empty((cin * k * k, cout), device=w.device, dtype=torch.float32)
    wsq = torch.empty((cin, cout), device=w.device, dtype=torch.float32) if want_wsq else None
    with _lib.device_of(w):
        _lib.check(lib.sg2_modconv2d_prep(wt.data_ptr(), _lib.ptr(wsq), w.data_ptr(), cin, cout, k,
                                          float(conv_scale), _lib.dtype_code(w), _lib.stream_of(w)),
                   "modconv2d_prep")
    return wt, wsq


def modulation(latent, mod_weight, mod_bias, wsq, cout, mod_scale, lr_mul, demodulate):
    """style [B,Cin] fp32 (+ demod [B,Cout] fp32) from latent [B, style_dim] (model.py:235-240)."""
    lib = _lib.load()
    lat = latent.contiguous()
    mw = mod_weight.detach().to(lat.dtype).contiguous()
    mb = mod_bias.detach().to(lat.dtype).contiguous()
    B, cin = lat.shape[0], mw.shape[0]
    style = torch.empty((B, cin), device=lat.device, dtype=torch.float32)
    demod = torch.empty((B, cout), device=lat.device, dtype=torch.float32) if demodulate else None
    with _lib.device_of(lat):
        _lib.check(lib.sg2_modulation_fwd(style.data_ptr(), _lib.ptr(demod), lat.data_ptr(), lat.stride(0),
                                          mw.data_ptr(), mb.data_ptr(), _lib.ptr(wsq), B, lat.shape[1], cin,
                                          cout, float(mod_scale), float(lr_mul), _lib.dtype_code(lat),
                                          _lib.stream_of(lat)), "modulation_fwd")
    return style, demod


def rgb_modconv(x, style, w2):
    """ToRGB's modulated 1x1 convolution without demodulation (model.py:350-355), no autograd: y[b,k] = sum_c w2[k,c] *
    style[b,c] * x[b,c] in one pass over the activation (csrc/torgb.cu, fp32 math); w2 [K,C] with the conv scale folded in."""
    _lib.require_cuda(x)
    x = x.contiguous()
    B, C, H, W = x.shape
    Kn = w2.shape[0]
    wf, sf = w2.detach().float().contiguous(), style.detach().float().contiguous()
    y = torch.empty((B, Kn, H, W), device=x.device, dtype=x.dtype)
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_rgb_modconv_fwd(y.data_ptr(), x.data_ptr(), wf.data_ptr(), sf.data_ptr(), B, C, Kn, H * W,
                                                   _lib.dtype_code(x), _lib.stream_of(x)), "rgb_modconv_fwd")
    return y


def conv_out_hw(h, w, k, mode):
    if mode == 0:
        return h, w
    if mode == 1:
        return (h - 1) * 2 + k, (w - 1) * 2 + k
    return (h - k) // 2 + 1, (w - k) // 2 + 1


def shared_conv(x, wt, style, demod, cout, k, mode):
    """out[b,co] = demod[b,co] * conv(x[b] * style[b], wt); mode 0 same / 1 transposed s2 / 2 stride 2."""
    _lib.require_cuda(x)
    lib = _lib.load()
    x = x.contiguous()
    B, cin, h, w = x.shape
    oh, ow = conv_out_hw(h, w, k, mode)
    out = torch.empty((B, cout, oh, ow), device=x.device, dtype=x.dtype)
    with _lib.device_of(x):
        _lib.check(lib.sg2_modconv2d_fwd(out.data_ptr(), x.data_ptr(), wt.data_ptr(), _lib.ptr(style),
                                         _lib.ptr(demod), B, cin, cout, h, w, k, mode, _lib.dtype_code(x),
                                         _lib.stream_of(x)), "modconv2d_fwd")
    return out


def noise_bias_act(x, noise, noise_weight, bias, act, alpha=0.2, act_scale=2 ** 0.5):
    """act(x + noise_weight*noise + bias[c]) * act_scale in one pass (model.py:282-287,335)."""
    _lib.require_cuda(x)
    lib = _lib.load()
    x = x.contiguous()
    B, Cn = x.shape[0], x.shape[1]
    HW = int(math.prod(x.shape[2:]))
    nstride = 0
    if noise is not None:
        noise = noise.detach().to(x.dtype).contiguous()
        if noise.numel() == B * HW:
            nstride = HW
        elif noise.numel() == HW:
            nstride = 0
        else:
            raise RuntimeError(f"noise of shape {tuple(noise.shape)} does not broadcast to {tuple(x.shape)}")
        noise_weight = noise_weight.detach().to(x.dtype).contiguous()
    if bias is not None:
        bias = bias.detach().to(x.dtype).contiguous()
    out = torch.empty_like(x)
    with _lib.device_of(x):
        _lib.check(lib.sg2_noise_bias_act(out.data_ptr(), x.data_ptr(), _lib.ptr(noise), nstride,
                                          _lib.ptr(noise_weight) if noise is not None else None, _lib.ptr(bias),
                                          B, Cn, HW, act, float(alpha), float(act_scale), _lib.dtype_code(x),
                                          _lib.stream_of(x)), "noise_bias_act")
    return out


def torgb_combine(conv, bias, skip, kernel, pad):
    """conv + bias[c] + upfirdn2d(skip, kernel, up=2, pad) in one pass (model.py:350-359)."""
    lib = _lib.load()
    conv = conv.contiguous()
    B, Cn, H, W = conv.shape
    b = None if bias is None else bias.detach().to(conv.dtype).reshape(-1).contiguous()
    taps = None
    kh = kw = 0
    if skip is not None:
        skip = skip.to(conv.dtype).contiguous()
        taps = kernel.detach().to(device=conv.device, dtype=torch.float32).contiguous()
        kh, kw = taps.shape
    out = torch.empty_like(conv)
    with _lib.device_of(conv):
        _lib.check(lib.sg2_torgb_combine(out.data_ptr(), conv.data_ptr(), _lib.ptr(b), _lib.ptr(skip),
                                         _lib.ptr(taps), kh, kw, pad[0], pad[1], B, Cn, H, W,
                                         _lib.dtype_code(conv), _lib.stream_of(conv)), "torgb_combine")
    return out


from .tc_route import (  # noqa: E402,F401  -- the tensor-core route, re-exported
    tc_grad_enabled, tc_grad, tc_conv_ok, _tc_pack, to_nhwc, to_nchw, _ones, tc_conv3x3_nhwc,
    tc_conv_transpose3x3_planes, planes_to_nchw, to_planes, tc_conv_transpose3x3_dgrad_planes, parts_to_nchw,
    tc_conv3x3, tc_conv_transpose3x3, tc_conv_transpose3x3_dgrad, _wgrad, ModulatedConvTCFunction, to_nchw_act,
    to_nhwc_actgrad, StyledConvTCFunction, cached_tc_packs, NoiseBiasActFunction, RgbModConvFunction)


class SharedConvFunction(torch.autograd.Function):
    """y = conv(x, W) with one weight tensor shared by the batch (the contraction inside
    ModulatedConv2d once modulation/demodulation are factored out).  forward and grad_x run on the
    sg2 kernels (fp32 SIMT; the tcgen05 kernel for 3x3 stride-1 layers when bf16 operands are allowed,
    see tc_grad); grad_W (only when the decoder itself is trained) uses the library wgrad."""

    @staticmethod
    def forward(ctx, x, weight4, mode):
        cout, cin, k, _ = weight4.shape
        ctx.save_for_backward(x, weight4)
        ctx.mode = mode
        ctx.tc = tc_conv_ok(x, weight4, mode)
        if ctx.tc:
            return tc_conv3x3(x, weight4) if mode == 0 else tc_conv_transpose3x3(x, weight4)
        wt, _ = conv_prep(weight4, 1.0, want_wsq=False)
        return shared_conv(x, wt, None, None, cout, k, mode)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, weight4 = ctx.saved_tensors
        mode = ctx.mode
        cout, cin, k, _ = weight4.shape
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if mode == 0:      # adjoint of a stride-1 'same' correlation: swap channels, flip taps
                wadj = weight4.detach().flip([2, 3]).transpose(0, 1)
                amode = 0
                if ctx.tc:
                    return tc_conv3x3(gy, wadj), (_wgrad(mode, x, weight4, gy) if ctx.needs_input_grad[1] else None), None
            elif mode == 1:    # adjoint of conv_transpose(stride 2) is conv(stride 2) with the same taps
                wadj = weight4.detach().transpose(0, 1)
                amode = 2
                if ctx.tc:
                    return tc_conv_transpose3x3_dgrad(gy, weight4), (_wgrad(mode, x, weight4, gy) if ctx.needs_input_grad[1] else None), None
            else:              # adjoint of conv(stride 2) is conv_transpose(stride 2)
                wadj = weight4.detach().transpose(0, 1)
                amode = 1
            wt, _ = conv_prep(wadj.contiguous(), 1.0, want_wsq=False)
            gx = shared_conv(gy.contiguous(), wt, None, None, cin, k, amode)
            if gx.shape[2:] != x.shape[2:]:   # stride-2 conv drops a trailing row/col when (H-k) is odd
                full = gx.new_zeros(x.shape)
                full[:, :, :gx.shape[2], :gx.shape[3]] = gx
                gx = full
        if ctx.needs_input_grad[1]:
            gw = _wgrad(mode, x, weight4, gy)
        return gx, gw, None
