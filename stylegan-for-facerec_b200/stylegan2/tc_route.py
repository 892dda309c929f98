"""Tensor-core route of the differentiable path: the autograd `Function`s and host wrappers that
put the decoder's convolutions -- forward and input gradient -- on the tcgen05 kernel when bf16 operands are allowed, with
modulation, demodulation, noise, bias, activation and their adjoint reductions folded into the two layout passes around
each convolution (csrc/synth.cu single-layer entries, csrc/layout_ops.cu, csrc/torgb.cu).

Re-exported by `functional` (model.py and stylegan2_ada reach everything as `K.<name>`).
"""
import ctypes as C
import os
import threading

import torch

from .. import _lib


# ---- tensor-core route of the differentiable path ---------------------------------------------------------------
# With bf16 operands allowed (Generator.precision == 'bf16', SG2_B200_PRECISION=bf16, or bfloat16 tensors) the 3x3
# stride-1 convolutions of the autograd path -- forward and input gradient, 2/3 of the decoder's training FLOPs --
# run on the tcgen05 kernel of the engine (sg2_conv3x3_tc: NHWC bf16 operands, fp32 accumulation).
_tc_grad = threading.local()


def tc_grad_enabled() -> bool:
    v = getattr(_tc_grad, "on", None)
    return (os.environ.get("SG2_B200_PRECISION", "auto") == "bf16") if v is None else v


class tc_grad:
    """context manager: allow (or forbid) bf16 tensor-core convolutions inside the differentiable path"""

    def __init__(self, on: bool):
        self.on = on

    def __enter__(self):
        self.prev = getattr(_tc_grad, "on", None)
        if self.on is not None:                 # None: leave the ambient setting (environment default)
            _tc_grad.on = self.on

    def __exit__(self, *exc):
        _tc_grad.on = self.prev
        return False


def tc_conv_ok(x, weight4, mode) -> bool:
    cout, cin, k, _ = weight4.shape
    return (mode in (0, 1) and k == 3 and x.dim() == 4 and x.shape[2] == x.shape[3] and x.shape[2] >= 4 and cin % 32 == 0
            and cout % 32 == 0 and x.shape[0] >= 1 and (x.dtype == torch.bfloat16 or tc_grad_enabled()))


def _tc_pack(weight4):
    """weight fp32 [Cout,Cin,3,3] -> bf16 [9][Cout][Cin] as the tensor-core kernel reads it"""
    cout, cin = weight4.shape[:2]
    w = weight4.detach().float().contiguous()
    wp = torch.empty((9, cout, cin), device=w.device, dtype=torch.bfloat16)
    with _lib.device_of(w):
        _lib.check(_lib.load().sg2_conv3x3_tc_pack(wp.data_ptr(), w.data_ptr(), cin, cout, 1.0, _lib.stream_of(w)),
                   "conv3x3_tc_pack")
    return wp


def to_nhwc(x, scale=None, other=None):
    """xh[b,y,x,c] = bf16(x[b,c,y,x] * scale[b,c]) in one pass (csrc/layout_ops.cu); with `other` (NHWC bf16, same
    shape as the result) also returns red[b,c] = sum_p x[b,c,p] * other[b,p,c], else None."""
    _lib.require_cuda(x)
    x = x.detach().contiguous()
    B, Cn, H, W = x.shape
    out = torch.empty((B, H, W, Cn), device=x.device, dtype=torch.bfloat16)
    red = torch.zeros((B, Cn), device=x.device, dtype=torch.float32) if other is not None else None
    sc = None if scale is None else scale.detach().float().contiguous()
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_nchw_to_nhwc_bf16(out.data_ptr(), x.data_ptr(), _lib.ptr(sc), _lib.ptr(other), _lib.ptr(red),
                                                     B, Cn, H * W, _lib.dtype_code(x), _lib.stream_of(x)), "nchw_to_nhwc_bf16")
    return out, red


def to_nchw(h, scale, dtype, other=None):
    """y[b,c,y,x] = scale[b,c] * h[b,y,x,c] (h NHWC bf16) in `dtype`; with `other` (NCHW, `dtype`) also returns
    red[b,c] = sum_p other[b,c,p] * h[b,p,c], else None."""
    B, H, W, Cn = h.shape
    out = torch.empty((B, Cn, H, W), device=h.device, dtype=dtype)
    red = torch.zeros((B, Cn), device=h.device, dtype=torch.float32) if other is not None else None
    sc = None if scale is None else scale.detach().float().contiguous()
    if other is not None:
        other = other.detach().to(dtype).contiguous()
    with _lib.device_of(h):
        _lib.check(_lib.load().sg2_nhwc_bf16_to_nchw(out.data_ptr(), h.data_ptr(), _lib.ptr(sc), _lib.ptr(other), _lib.ptr(red),
                                                     B, Cn, H * W, _lib.dtype_code(out), _lib.stream_of(h)), "nhwc_bf16_to_nchw")
    return out, red


def _ones(B, n, device):
    return torch.ones((B, n), device=device, dtype=torch.float32)


def tc_conv3x3_nhwc(xh, wp, scale=None):
    """NHWC bf16 [B,r,r,Cin] -> [B,r,r,Cout]: scale[b,co] * conv3x3 'same' with packed weights wp [9,Cout,Cin]"""
    B, r, _, cin = xh.shape
    cout = wp.shape[1]
    sc = _ones(B, cout, xh.device) if scale is None else scale.detach().float().contiguous()
    out = torch.empty((B, r, r, cout), device=xh.device, dtype=torch.bfloat16)
    with _lib.device_of(xh):
        _lib.check(_lib.load().sg2_conv3x3_tc(out.data_ptr(), xh.data_ptr(), wp.data_ptr(), sc.data_ptr(), B, r, cin, cout,
                                              _lib.stream_of(xh)), "conv3x3_tc")
    return out


def tc_conv_transpose3x3_planes(xh, wp):
    """NHWC bf16 [B,r,r,Cin] -> conv_transpose2d(stride 2) as its four polyphase planes [4,B,r+1,r+1,Cout] bf16
    (pixel (y, x) of the (2r+1)^2 result lives in plane (y&1)*2 + (x&1) at (y>>1, x>>1); the last row / column of the odd
    planes is not written)."""
    B, r, _, cin = xh.shape
    cout = wp.shape[1]
    planes = torch.empty((4, B, r + 1, r + 1, cout), device=xh.device, dtype=torch.bfloat16)
    with _lib.device_of(xh):
        _lib.check(_lib.load().sg2_conv_transpose3x3_tc(planes.data_ptr(), xh.data_ptr(), wp.data_ptr(),
                                                        _ones(B, cout, xh.device).data_ptr(), B, r, cin, cout,
                                                        _lib.stream_of(xh)), "conv_transpose3x3_tc")
    return planes


def planes_to_nchw(planes, scale, dtype, other=None):
    """polyphase planes [4,B,P,P,C] bf16 -> [B,C,2P-1,2P-1] `dtype`, times scale[b,c]; with `other` (NCHW) also
    red[b,c] = sum_p other * planes (csrc/layout_ops.cu, polyphase addressing: no interleaving copy)"""
    _, B, P, _, Cn = planes.shape
    R = 2 * P - 1
    out = torch.empty((B, Cn, R, R), device=planes.device, dtype=dtype)
    red = torch.zeros((B, Cn), device=planes.device, dtype=torch.float32) if other is not None else None
    sc = None if scale is None else scale.detach().float().contiguous()
    if other is not None:
        other = other.detach().to(dtype).contiguous()
    with _lib.device_of(planes):
        _lib.check(_lib.load().sg2_polyphase_bf16_to_nchw(out.data_ptr(), planes.data_ptr(), _lib.ptr(sc), _lib.ptr(other),
                                                          _lib.ptr(red), B, Cn, R, _lib.dtype_code(out), _lib.stream_of(planes)),
                   "polyphase_bf16_to_nchw")
    return out, red


def to_planes(x, scale=None, other=None):
    """[B,C,R,R] (R odd) -> zero-padded polyphase planes [4,B,P,P,C] bf16 of x * scale[b,c]; with `other` (planes of the
    same shape) also red[b,c] = sum_p x * other"""
    x = x.detach().contiguous()
    B, Cn, R, _ = x.shape
    P = (R + 1) // 2
    planes = torch.zeros((4, B, P, P, Cn), device=x.device, dtype=torch.bfloat16)
    red = torch.zeros((B, Cn), device=x.device, dtype=torch.float32) if other is not None else None
    sc = None if scale is None else scale.detach().float().contiguous()
    with _lib.device_of(x):
        _lib.check(_lib.load().sg2_nchw_to_polyphase_bf16(planes.data_ptr(), x.data_ptr(), _lib.ptr(sc), _lib.ptr(other),
                                                          _lib.ptr(red), B, Cn, R, _lib.dtype_code(x), _lib.stream_of(x)),
                   "nchw_to_polyphase_bf16")
    return planes, red


def tc_conv_transpose3x3_dgrad_planes(planes, wp_t):
    """input gradient of the transposed conv: gx[i,j,ci] = sum_{a,b,co} gy[2i+a, 2j+b, co] * w[co,ci,a,b], a stride-2
    convolution evaluated as the sum of four stride-1 convolutions (4, 2, 2, 1 taps) over the zero-padded polyphase
    planes of gy [4,B,r+1,r+1,Cout]; wp_t = pack(w^T) [9,Cin,Cout] -> the four components [4,B,r+1,r+1,Cin] bf16, whose
    top-left r x r corners sum to the gradient (parts_to_nchw)."""
    lib = _lib.load()
    _, B, P, _, cout = planes.shape
    cin = wp_t.shape[1]
    r = P - 1
    ones = _ones(B, cin, planes.device)
    parts = torch.empty((4, B, P, P, cin), device=planes.device, dtype=torch.bfloat16)
    with _lib.device_of(planes):
        st = _lib.stream_of(planes)
        for s in range(4):
            py, px = s >> 1, s & 1
            taps = [(da, db, (2 * da + py) * 3 + 2 * db + px) for da in range(2 - py) for db in range(2 - px)]
            flat = (C.c_int * (3 * len(taps)))(*[v for t in taps for v in t])
            _lib.check(lib.sg2_conv_taps_tc(parts[s].data_ptr(), planes[s].data_ptr(), wp_t.data_ptr(), ones.data_ptr(), B, P, cout,
                                            cin, flat, len(taps), st), "conv_taps_tc")
    return parts


def parts_to_nchw(parts, scale, dtype, other=None):
    """the four polyphase components [4,B,P,P,C] of the transposed conv's input gradient -> scale[b,c] * their fp32 sum
    over the valid (P-1)^2 corner as [B,C,P-1,P-1] `dtype`; with `other` (NCHW) also red[b,c] = sum_p other * sum"""
    n, B, P, _, Cn = parts.shape
    r = P - 1
    out = torch.empty((B, Cn, r, r), device=parts.device, dtype=dtype)
    red = torch.zeros((B, Cn), device=parts.device, dtype=torch.float32) if other is not None else None
    sc = None if scale is None else scale.detach().float().contiguous()
    if other is not None:
        other = other.detach().to(dtype).contiguous()
    with _lib.device_of(parts):
        _lib.check(_lib.load().sg2_sum_parts_bf16_to_nchw(out.data_ptr(), parts.data_ptr(), n, P, _lib.ptr(sc), _lib.ptr(other),
                                                          _lib.ptr(red), B, Cn, r, _lib.dtype_code(out), _lib.stream_of(parts)),
                   "sum_parts_bf16_to_nchw")
    return out, red


def tc_conv3x3(x, weight4, scale=None):
    """y = scale[b,co] * conv2d(x, weight4, padding=1) on the tensor-core kernel: x [B,Cin,r,r] any float dtype -> same dtype."""
    xh, _ = to_nhwc(x)
    return to_nchw(tc_conv3x3_nhwc(xh, _tc_pack(weight4), scale), None, x.dtype)[0]


def tc_conv_transpose3x3(x, weight4):
    """y = conv_transpose2d(x, weight4^T, stride=2) (the up-sampling ModulatedConv2d before its blur) on the tensor-core
    kernel: x [B,Cin,r,r], weight4 [Cout,Cin,3,3] -> [B,Cout,2r+1,2r+1] in x.dtype."""
    xh, _ = to_nhwc(x)
    return planes_to_nchw(tc_conv_transpose3x3_planes(xh, _tc_pack(weight4)), None, x.dtype)[0]


def tc_conv_transpose3x3_dgrad(gy, weight4):
    planes, _ = to_planes(gy)
    return parts_to_nchw(tc_conv_transpose3x3_dgrad_planes(planes, _tc_pack(weight4.detach().transpose(0, 1))), None, gy.dtype)[0]


def _wgrad(mode, x, weight4, gy):
    """weight gradient of the shared-weight convolution (library wgrad, true fp32)"""
    import torch.nn.grad as G
    cout, cin, k, _ = weight4.shape
    xf, gf = x.detach().float(), gy.detach().float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        if mode == 0:
            gw = G.conv2d_weight(xf, weight4.shape, gf, padding=k // 2)
        elif mode == 2:
            gw = G.conv2d_weight(xf, weight4.shape, gf, stride=2)
        else:   # y = conv_transpose(x, W^T): dW[co,ci] = corr(gy[co], x[ci]) at stride 2
            gw = G.conv2d_weight(gf, (cin, cout, k, k), xf, stride=2).transpose(0, 1)
    return gw.to(weight4.dtype)


class ModulatedConvTCFunction(torch.autograd.Function):
    """y = d[b,co] * conv(W, s[b,ci] * x) -- the factored ModulatedConv2d (model.py:232-273) -- on the tensor-core kernel,
    mode 0 (3x3 'same') or 1 (stride-2 transposed).  Three passes forward (modulate + to NHWC bf16, conv, demodulate + to
    NCHW) and three backward; the adjoint passes also produce grad_s and grad_d, so nothing else touches the activations.
    x [B,Cin,r,r]; s [B,Cin], d [B,Cout] fp32 (d may be None); weight4 [Cout,Cin,3,3]."""

    @staticmethod
    def forward(ctx, x, s, d, weight4, mode, wp=None, wp_adj=None):
        xh, _ = to_nhwc(x, s)
        wp = _tc_pack(weight4) if wp is None else wp
        ctx.wp_adj = wp_adj
        if mode == 0:
            yh = tc_conv3x3_nhwc(xh, wp)
            y, _ = to_nchw(yh, d, x.dtype)
        else:                                   # the conv output stays in its polyphase planes
            yh = tc_conv_transpose3x3_planes(xh, wp)
            y, _ = planes_to_nchw(yh, d, x.dtype)
        ctx.save_for_backward(x, s, d, weight4, yh)
        ctx.mode = mode
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, s, d, weight4, yh = ctx.saved_tensors
        mode = ctx.mode
        gy = gy.contiguous()
        need_d = d is not None and ctx.needs_input_grad[2]
        # gh = bf16(d * gy) (mode 1: straight into zero-padded polyphase planes), gd = sum_p gy * conv
        gh, gd = (to_nhwc if mode == 0 else to_planes)(gy, d, other=yh if need_d else None)
        gx = gs = gw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            xo = x if ctx.needs_input_grad[1] else None
            if mode == 0:                                                          # gx = s * g, gs = sum_p x * g
                gxh = tc_conv3x3_nhwc(gh, ctx.wp_adj if ctx.wp_adj is not None else _tc_pack(weight4.detach().flip([2, 3]).transpose(0, 1)))
                gx, gs = to_nchw(gxh, s, x.dtype, other=xo)
            else:
                parts = tc_conv_transpose3x3_dgrad_planes(gh, ctx.wp_adj if ctx.wp_adj is not None else _tc_pack(weight4.detach().transpose(0, 1)))
                gx, gs = parts_to_nchw(parts, s, x.dtype, other=xo)
            if gs is not None:
                gs = gs.to(s.dtype)
        if ctx.needs_input_grad[3]:
            B = x.shape[0]
            xm = x.detach().float() * s.detach().float().view(B, -1, 1, 1)
            gc = gy.float() if d is None else gy.float() * d.detach().float().view(B, -1, 1, 1)
            gw = _wgrad(mode, xm, weight4, gc)
        if gd is not None:
            gd = gd.to(d.dtype)
        return gx, gs, gd, gw, None, None, None


def to_nchw_act(h, scale, dtype, noise, noise_weight, bias, alpha, gain):
    """y = lrelu(scale[b,c] * h[b,p,c] + noise_weight * noise[b or 0, p] + bias[c], alpha) * gain, NHWC bf16 -> NCHW `dtype`
    in one pass (demodulation + NoiseInjection + FusedLeakyReLU, model.py:239-240,282-287,335)"""
    B, H, W, Cn = h.shape
    out = torch.empty((B, Cn, H, W), device=h.device, dtype=dtype)
    sc = None if scale is None else scale.detach().float().contiguous()
    nstride = 0
    if noise is not None:
        noise = noise.detach().to(dtype).contiguous()
        if noise.numel() == B * H * W:
            nstride = H * W
        elif noise.numel() != H * W:
            raise RuntimeError(f"noise of shape {tuple(noise.shape)} does not broadcast to [{B}, 1, {H}, {W}]")
        noise_weight = noise_weight.detach().to(dtype).contiguous()
    if bias is not None:
        bias = bias.detach().to(dtype).contiguous()
    with _lib.device_of(h):
        _lib.check(_lib.load().sg2_nhwc_bf16_to_nchw_act(out.data_ptr(), h.data_ptr(), _lib.ptr(sc), _lib.ptr(noise), nstride,
                                                         _lib.ptr(noise_weight) if noise is not None else None, _lib.ptr(bias),
                                                         float(alpha), float(gain), B, Cn, H * W, _lib.dtype_code(out),
                                                         _lib.stream_of(h)), "nhwc_bf16_to_nchw_act")
    return out


def to_nhwc_actgrad(gy, y, alpha, gain, scale, other=None, want_sum=False):
    """adjoint of to_nchw_act: g = gy * (y > 0 ? 1 : alpha) * gain; returns (bf16(g * scale) as NHWC,
    red[b,c] = sum_p g * other[b,p,c] or None, sum[b,c] = sum_p g or None)"""
    gy = gy.detach().contiguous()
    y = y.detach().to(gy.dtype).contiguous()
    B, Cn, H, W = gy.shape
    out = torch.empty((B, H, W, Cn), device=gy.device, dtype=torch.bfloat16)
    red = torch.zeros((B, Cn), device=gy.device, dtype=torch.float32) if other is not None else None
    tot = torch.zeros((B, Cn), device=gy.device, dtype=torch.float32) if want_sum else None
    sc = None if scale is None else scale.detach().float().contiguous()
    with _lib.device_of(gy):
        _lib.check(_lib.load().sg2_nchw_to_nhwc_bf16_actgrad(out.data_ptr(), gy.data_ptr(), y.data_ptr(), float(alpha), float(gain),
                                                             _lib.ptr(sc), _lib.ptr(other), _lib.ptr(red), _lib.ptr(tot), B, Cn,
                                                             H * W, _lib.dtype_code(gy), _lib.stream_of(gy)),
                   "nchw_to_nhwc_bf16_actgrad")
    return out, red, tot


class StyledConvTCFunction(torch.autograd.Function):
    """out = lrelu(d * conv3x3(W, s * x) + noise_weight * noise + bias, alpha) * gain -- a whole non-resampling StyledConv
    (model.py:331-337) in three passes each way: modulate + to NHWC bf16, tensor-core conv, demodulate + noise + bias +
    activation + to NCHW; backward: activation gradient + demodulation (+ grad_d, grad_bias) + to NHWC, conv with the
    adjoint weights, style (+ grad_s) + to NCHW."""

    @staticmethod
    def forward(ctx, x, s, d, weight4, noise, noise_weight, bias, alpha, gain, wp=None, wp_adj=None):
        xh, _ = to_nhwc(x, s)
        yh = tc_conv3x3_nhwc(xh, _tc_pack(weight4) if wp is None else wp)
        ctx.wp_adj = wp_adj
        out = to_nchw_act(yh, d, x.dtype, noise, noise_weight, bias, alpha, gain)
        ctx.save_for_backward(x, s, d, weight4, yh, out, noise, noise_weight)
        ctx.act = (alpha, gain)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        x, s, d, weight4, yh, out, noise, noise_weight = ctx.saved_tensors
        alpha, gain = ctx.act
        need = ctx.needs_input_grad
        need_d = d is not None and need[2]
        gh, gd, tot = to_nhwc_actgrad(gout, out, alpha, gain, d, other=yh if need_d else None, want_sum=need[6])
        gx = gs = gw = gnoise = gnw = gbias = None
        if need[0] or need[1]:
            gxh = tc_conv3x3_nhwc(gh, ctx.wp_adj if ctx.wp_adj is not None else _tc_pack(weight4.detach().flip([2, 3]).transpose(0, 1)))
            gx, gs = to_nchw(gxh, s, x.dtype, other=x if need[1] else None)
            if gs is not None:
                gs = gs.to(s.dtype)
        if gd is not None:
            gd = gd.to(d.dtype)
        if need[6]:
            gbias = tot.sum(0).to(out.dtype)
        if need[3] or (noise is not None and (need[4] or need[5])):
            # rare (the decoder itself trains): the activation gradient as a tensor, from the op kernel
            from .op.fused_act import bias_act
            g = bias_act(gout.contiguous(), None, out, 3, 1, alpha, gain)
            B = x.shape[0]
            if noise is not None and (need[4] or need[5]):
                t = g.float().sum(1, keepdim=True)                                           # [B,1,H,W]
                if need[4]:
                    gn = noise_weight.detach().float().view(1, 1, 1, 1) * t
                    gnoise = (gn if noise.shape[0] == B else gn.sum(0, keepdim=True)).reshape(noise.shape).to(noise.dtype)
                if need[5]:
                    gnw = (t * noise.detach().float().reshape(-1, 1, *t.shape[2:])).sum().reshape(noise_weight.shape).to(noise_weight.dtype)
            if need[3]:
                xm = x.detach().float() * s.detach().float().view(B, -1, 1, 1)
                gc = g.float() if d is None else g.float() * d.detach().float().view(B, -1, 1, 1)
                gw = _wgrad(0, xm, weight4, gc)
        return gx, gs, gd, gw, gnoise, gnw, gbias, None, None, None, None


def cached_tc_packs(owner, weight4_fn, version_key, mode):
    """(w4, wp, wp_adj, wsq) of a FROZEN conv weight for the tensor-core route, cached on `owner` until the parameter
    changes: w4 = weight4_fn() detached [Cout,Cin,3,3], wp / wp_adj its packed forward / adjoint forms, wsq [Cout,Cin] the
    per-(co,ci) sum of squared taps the demodulation needs.  Saves a scale, two flips/transposes and two pack launches
    per layer per step, forward and backward."""
    c = owner.__dict__.get("_tc_cache")
    if c is None or c[0] != version_key:
        with torch.no_grad():
            w4 = weight4_fn().detach()
            adj = w4.flip([2, 3]).transpose(0, 1) if mode == 0 else w4.transpose(0, 1)
            c = (version_key, w4, _tc_pack(w4), _tc_pack(adj), w4.float().pow(2).sum([2, 3]))
        owner.__dict__["_tc_cache"] = c
    return c[1:]


class NoiseBiasActFunction(torch.autograd.Function):
    """out = lrelu(x + noise_weight * noise + bias[c], alpha) * gain (NoiseInjection + FusedLeakyReLU, model.py:282-287,335)
    as one pass each way (first-order autograd); used after the blur of the up-sampling layers on the tensor-core route."""

    @staticmethod
    def forward(ctx, x, noise, noise_weight, bias, alpha, gain):
        from .functional import noise_bias_act
        out = noise_bias_act(x, noise, noise_weight, bias, act=3, alpha=alpha, act_scale=gain)
        ctx.save_for_backward(out, noise, noise_weight)
        ctx.act = (alpha, gain)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        from .op.fused_act import bias_act, grad_bias_reduce
        out, noise, noise_weight = ctx.saved_tensors
        alpha, gain = ctx.act
        need = ctx.needs_input_grad
        gx = bias_act(g.contiguous(), None, out, 3, 1, alpha, gain)
        gnoise = gnw = gbias = None
        if need[3]:
            gbias = grad_bias_reduce(gx)
        if noise is not None and (need[1] or need[2]):
            t = gx.float().sum(1, keepdim=True)                                              # [B,1,H,W]
            if need[1]:
                gn = noise_weight.detach().float().view(1, 1, 1, 1) * t
                gnoise = (gn if noise.shape[0] == gx.shape[0] else gn.sum(0, keepdim=True)).reshape(noise.shape).to(noise.dtype)
            if need[2]:
                gnw = (t * noise.detach().float().reshape(-1, 1, *t.shape[2:])).sum().reshape(noise_weight.shape).to(noise_weight.dtype)
        return gx, gnoise, gnw, gbias, None, None


class RgbModConvFunction(torch.autograd.Function):
    """y[b,k] = sum_c w[k,c] * s[b,c] * x[b,c]: the modulated 1x1 convolution of ToRGB (model.py:350-355, no
    demodulation) -- one pass over the activation each way (csrc/torgb.cu), fp32 math.  x [B,C,H,W]; s [B,C];
    w2 [K,C] with the conv scale folded in."""

    @staticmethod
    def forward(ctx, x, s, w2):
        _lib.require_cuda(x)
        x = x.contiguous()
        B, Cn, H, W = x.shape
        K_ = w2.shape[0]
        wf, sf = w2.detach().float().contiguous(), s.detach().float().contiguous()
        y = torch.empty((B, K_, H, W), device=x.device, dtype=x.dtype)
        with _lib.device_of(x):
            _lib.check(_lib.load().sg2_rgb_modconv_fwd(y.data_ptr(), x.data_ptr(), wf.data_ptr(), sf.data_ptr(), B, Cn, K_, H * W,
                                                       _lib.dtype_code(x), _lib.stream_of(x)), "rgb_modconv_fwd")
        ctx.save_for_backward(x, s, w2)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, s, w2 = ctx.saved_tensors
        B, Cn, H, W = x.shape
        K_ = w2.shape[0]
        gy = gy.contiguous().to(x.dtype)
        gx = gs = gw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            wf, sf = w2.detach().float().contiguous(), s.detach().float().contiguous()
            gx = torch.empty_like(x)
            gs = torch.zeros((B, Cn), device=x.device, dtype=torch.float32)
            with _lib.device_of(x):
                _lib.check(_lib.load().sg2_rgb_modconv_bwd(gx.data_ptr(), gs.data_ptr(), gy.data_ptr(), x.data_ptr(), wf.data_ptr(),
                                                           sf.data_ptr(), B, Cn, K_, H * W, _lib.dtype_code(x), _lib.stream_of(x)),
                           "rgb_modconv_bwd")
            gs = gs.to(s.dtype)
        if ctx.needs_input_grad[2]:            # rare (the decoder itself trains): gw[k,c] = sum_{b,p} gy[b,k,p] * s[b,c] * x[b,c,p]
            gw = torch.einsum("bkp,bcp->bkc", gy.float().flatten(2), x.detach().float().flatten(2))
            gw = (gw * s.detach().float().unsqueeze(1)).sum(0).to(w2.dtype)
        return gx, gs, gw
