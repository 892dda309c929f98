"""GPU: the stand-alone tensor-core 3x3 convolution (sg2_conv3x3_tc) and the bf16 route of the differentiable path
vs fp64 F.conv2d on CPU (the call ModulatedConv2d.forward makes, model.py:269-273 of the reference)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _K():
    return importlib.import_module("stylegan-for-facerec_b200.stylegan2.functional")


def _bf(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("B,cin,cout,r", [(2, 64, 32, 16), (3, 512, 512, 4), (1, 32, 32, 64), (2, 128, 256, 32), (5, 64, 64, 8),
                                          (1, 256, 256, 32), (2, 96, 160, 12), (1, 64, 64, 20), (4, 32, 64, 5)])
def test_conv3x3_tc_vs_fp64(sg2, B, cin, cout, r):
    K = _K()
    g = torch.Generator().manual_seed(B * 1000 + cin + r)
    x = _bf(torch.randn(B, cin, r, r, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
    sc = torch.rand(B, cout, generator=g) + 0.5
    ref = F.conv2d(x.double(), w.double(), padding=1) * sc.double().view(B, cout, 1, 1)
    y = K.tc_conv3x3(x.to(DEV), w.to(DEV), sc.to(DEV))
    assert y.shape == ref.shape and y.dtype == torch.float32
    err = (y.cpu().double() - ref).abs().max().item()
    assert err <= 6e-3 * ref.abs().max().item(), (err, ref.abs().max().item())     # bf16 output rounding: 2^-8 relative
    yb = K.tc_conv3x3(x.to(DEV).bfloat16(), w.to(DEV))                            # bf16 in -> bf16 out, no scale
    ref1 = F.conv2d(x.double(), w.double(), padding=1)
    assert yb.dtype == torch.bfloat16 and (yb.cpu().double() - ref1).abs().max() <= 6e-3 * ref1.abs().max()


def test_conv3x3_tc_rejects_what_it_cannot_run(sg2):
    K = _K()
    x = torch.randn(1, 24, 8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="Cin"):
        K.tc_conv3x3(x, torch.randn(32, 24, 3, 3, device=DEV))
    assert not K.tc_conv_ok(x, torch.randn(32, 24, 3, 3), 0)
    xb = torch.randn(1, 32, 8, 8, device=DEV)
    w = torch.randn(32, 32, 3, 3)
    assert not K.tc_conv_ok(xb, w, 0)                                  # fp32 tensors, bf16 operands not allowed
    with K.tc_grad(True):
        assert K.tc_conv_ok(xb, w, 0) and K.tc_conv_ok(xb, w, 1) and not K.tc_conv_ok(xb, w, 2) and not K.tc_conv_ok(xb, w[:, :, :1, :1], 0)
    assert K.tc_conv_ok(xb.bfloat16(), w, 0)
    with K.tc_grad(False):
        assert not K.tc_conv_ok(xb, w, 0)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,cin,cout,r", [(2, 64, 32, 16), (3, 512, 512, 4), (1, 32, 64, 32), (2, 128, 64, 7)])
def test_shared_conv_function_tensor_core_gradients(sg2, B, cin, cout, r, mode):
    """mode 0: F.conv2d(x, w, padding=1) (model.py:269-273); mode 1: F.conv_transpose2d(x, w^T, stride=2) (model.py:246-252)"""
    K = _K()
    g = torch.Generator().manual_seed(cin + r)
    x = _bf(torch.randn(B, cin, r, r, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
    ro = r if mode == 0 else 2 * r + 1
    gy = _bf(torch.randn(B, cout, ro, ro, generator=g))
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv2d(x64, w64, padding=1) if mode == 0 else F.conv_transpose2d(x64, w64.transpose(0, 1), stride=2)
    ref.backward(gy.double())
    xd, wd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    with K.tc_grad(True):
        y = K.SharedConvFunction.apply(xd, wd, mode)
    y.backward(gy.to(DEV))
    for name, a, b in (("y", y.detach(), ref.detach()), ("gx", xd.grad, x64.grad), ("gw", wd.grad, w64.grad)):
        err = (a.cpu().double() - b).abs().max().item()
        assert a.shape == b.shape and err <= 6e-3 * b.abs().max().item(), (name, err, b.abs().max().item())
    # and the fp32 route is untouched when bf16 operands are not allowed
    xe = x.to(DEV).requires_grad_(True)
    ye = K.SharedConvFunction.apply(xe, w.to(DEV), mode)
    assert (ye.detach().cpu().double() - ref.detach()).abs().max() <= 1e-5 * ref.abs().max()


def test_conv_taps_tc_rejects_bad_taps(sg2):
    import ctypes as C
    lib = sg2._lib.load()
    x = torch.zeros(1, 8, 8, 32, device=DEV, dtype=torch.bfloat16)
    wp = torch.zeros(9, 32, 32, device=DEV, dtype=torch.bfloat16)
    sc = torch.ones(1, 32, device=DEV)
    out = torch.empty(1, 8, 8, 32, device=DEV, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for taps in ((2, 0, 0), (0, 0, 9), (0, -2, 1)):
        arr = (C.c_int * 3)(*taps)
        assert lib.sg2_conv_taps_tc(out.data_ptr(), x.data_ptr(), wp.data_ptr(), sc.data_ptr(), 1, 8, 32, 32, arr, 1, st) != 0
    assert lib.sg2_conv_taps_tc(out.data_ptr(), x.data_ptr(), wp.data_ptr(), sc.data_ptr(), 1, 8, 32, 32, None, 1, st) != 0
    arr = (C.c_int * 3)(0, 0, 4)                                       # the centre tap alone: a 1x1 convolution
    xr = torch.randn(1, 8, 8, 32, device=DEV).bfloat16()
    wr = torch.randn(9, 32, 32, device=DEV).bfloat16()
    assert lib.sg2_conv_taps_tc(out.data_ptr(), xr.data_ptr(), wr.data_ptr(), sc.data_ptr(), 1, 8, 32, 32, arr, 1, st) == 0
    ref = xr.float() @ wr[4].float().t()
    assert (out.float() - ref).abs().max() <= 6e-3 * ref.abs().max()


def test_generator_bf16_gradients_vs_oracle(sg2, oracle):
    """precision='bf16' with gradients: tensor-core convolutions inside the differentiable path"""
    sd = oracle.init_state_dict(16, 512, 2)
    G = sg2.Generator(16, 512, 2)
    G.load_state_dict(sd)
    G = G.to(DEV).eval()
    G.precision = 'bf16'
    for p in G.parameters():
        p.requires_grad_(False)
    lat = 0.5 * oracle.named_randn("grad:lat", (2, 6, 512), 3)
    gy = oracle.named_randn("grad:gy", (2, 3, 16, 16), 3)
    ld = lat.to(DEV).requires_grad_(True)
    img, _ = G([ld], input_is_latent=True, randomize_noise=False)
    img.backward(gy.to(DEV))
    lo = lat.double().requires_grad_(True)
    imgo, _ = oracle.generator_forward({k: v.double() for k, v in sd.items()}, 16, [lo], n_mlp=2, input_is_latent=True,
                                       randomize_noise=False)
    imgo.backward(gy.double())
    assert (img.detach().cpu().double() - imgo.detach()).abs().max() <= 3e-2 * imgo.abs().max()
    err = ld.grad.cpu().double() - lo.grad
    # bf16 operand rounding (2^-9 per element) through 5 conv layers forward and back, amplified by the cancellation
    # between the modulation and demodulation terms of d/ds: 4 % measured; a wrong tap flip or channel swap gives O(1)
    assert err.norm() <= 8e-2 * lo.grad.norm(), (err.norm() / lo.grad.norm()).item()
    G.precision = 'exact'
    ld2 = lat.to(DEV).requires_grad_(True)
    img2, _ = G([ld2], input_is_latent=True, randomize_noise=False)
    assert (img2.detach().cpu().double() - imgo.detach()).abs().max() <= 1e-3 * max(1.0, imgo.abs().max().item())


@pytest.mark.parametrize("B,Cn,H,W", [(2, 32, 5, 5), (1, 96, 33, 33), (3, 160, 8, 8), (1, 64, 64, 64), (2, 34, 7, 9)])
def test_layout_passes_vs_torch(sg2, B, Cn, H, W):
    """NCHW <-> NHWC bf16 with the per-(sample, channel) factor and the adjoint's reduction in the same pass"""
    K = _K()
    g = torch.Generator().manual_seed(Cn + H)
    x = torch.randn(B, Cn, H, W, generator=g)
    sc = torch.rand(B, Cn, generator=g) + 0.5
    oth = torch.randn(B, H, W, Cn, generator=g).bfloat16()
    for dt in (torch.float32, torch.bfloat16):
        xd = x.to(dt)
        ref = (xd.double() * sc.double().view(B, Cn, 1, 1)).permute(0, 2, 3, 1)
        h, red = K.to_nhwc(xd.to(DEV), sc.to(DEV), other=oth.to(DEV))
        assert h.dtype == torch.bfloat16 and h.shape == (B, H, W, Cn)
        assert (h.cpu().double() - ref).abs().max() <= 2 ** -8 * ref.abs().max()
        rref = (xd.double().permute(0, 2, 3, 1) * oth.double()).sum((1, 2))
        assert (red.cpu().double() - rref).abs().max() <= 1e-5 * rref.abs().max() + 1e-5
        h2, none = K.to_nhwc(xd.to(DEV))
        assert none is None and torch.equal(h2.cpu(), xd.permute(0, 2, 3, 1).bfloat16())
        # and back
        hb = torch.randn(B, H, W, Cn, generator=g).bfloat16()
        y, red = K.to_nchw(hb.to(DEV), sc.to(DEV), dt, other=xd.to(DEV))
        yref = hb.double().permute(0, 3, 1, 2) * sc.double().view(B, Cn, 1, 1)
        assert y.dtype == dt and (y.cpu().double() - yref).abs().max() <= (2 ** -8 if dt == torch.bfloat16 else 1e-6) * yref.abs().max()
        rref = (xd.double() * hb.double().permute(0, 3, 1, 2)).sum((2, 3))
        assert (red.cpu().double() - rref).abs().max() <= 1e-5 * rref.abs().max() + 1e-5
        y2, none = K.to_nchw(hb.to(DEV), None, dt)
        assert none is None and torch.equal(y2.cpu(), hb.permute(0, 3, 1, 2).to(dt))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,cin,cout,r,demod", [(2, 64, 32, 16, True), (3, 512, 512, 4, True), (2, 32, 64, 9, False)])
def test_modulated_conv_tc_function_gradients(sg2, B, cin, cout, r, demod, mode):
    """y = d * conv(W, s * x) in three passes each way: every gradient (x, s, d, W) vs fp64 autograd"""
    K = _K()
    g = torch.Generator().manual_seed(cin + r + mode)
    x = torch.randn(B, cin, r, r, generator=g)
    s = torch.rand(B, cin, generator=g) + 0.5
    d = torch.rand(B, cout, generator=g) + 0.5 if demod else None
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    ro = r if mode == 0 else 2 * r + 1
    gy = torch.randn(B, cout, ro, ro, generator=g)
    leaves = [t.double().requires_grad_(True) for t in (x, s, w)] + ([d.double().requires_grad_(True)] if demod else [])
    x64, s64, w64 = leaves[:3]
    xm = x64 * s64.view(B, cin, 1, 1)
    ref = F.conv2d(xm, w64, padding=1) if mode == 0 else F.conv_transpose2d(xm, w64.transpose(0, 1), stride=2)
    if demod:
        ref = ref * leaves[3].view(B, cout, 1, 1)
    ref.backward(gy.double())
    dev = [t.to(DEV).requires_grad_(True) for t in (x, s, w)] + ([d.to(DEV).requires_grad_(True)] if demod else [])
    y = K.ModulatedConvTCFunction.apply(dev[0], dev[1], dev[3] if demod else None, dev[2], mode)
    assert y.shape == ref.shape and (y.detach().cpu().double() - ref.detach()).abs().max() <= 1e-2 * ref.abs().max()
    y.backward(gy.to(DEV))
    for name, a, b in zip(("gx", "gs", "gw", "gd"), dev, leaves):
        err = (a.grad.cpu().double() - b.grad).abs().max().item()
        assert a.grad.shape == b.grad.shape and err <= 1.2e-2 * b.grad.abs().max().item(), (name, err, b.grad.abs().max().item())


@pytest.mark.parametrize("B,cin,cout,r,shared_noise", [(2, 64, 32, 16, False), (3, 512, 512, 4, True), (2, 32, 96, 9, False)])
def test_styled_conv_tc_function_gradients(sg2, B, cin, cout, r, shared_noise):
    """a whole non-resampling StyledConv (model.py:331-337) in three passes each way: output and every gradient vs fp64
    autograd; the upstream gradient is zeroed where the pre-activation is within bf16 rounding of the leaky-relu kink"""
    K = _K()
    g = torch.Generator().manual_seed(cin + r)
    x = torch.randn(B, cin, r, r, generator=g)
    s = torch.rand(B, cin, generator=g) + 0.5
    d = torch.rand(B, cout, generator=g) + 0.5
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    nz = torch.randn(1 if shared_noise else B, 1, r, r, generator=g)
    nw = torch.tensor([0.3])
    bias = 0.2 * torch.randn(cout, generator=g)
    leaves = [t.double().requires_grad_(True) for t in (x, s, d, w, nz, nw, bias)]
    x64, s64, d64, w64, nz64, nw64, b64 = leaves
    pre = F.conv2d(x64 * s64.view(B, cin, 1, 1), w64, padding=1) * d64.view(B, cout, 1, 1) + nw64 * nz64 + b64.view(1, cout, 1, 1)
    ref = F.leaky_relu(pre, 0.2) * 2 ** 0.5
    gy = torch.randn(B, cout, r, r, generator=g) * (pre.detach().abs() > 3e-2 * pre.detach().abs().max())
    ref.backward(gy.double())
    dev = [t.to(DEV).requires_grad_(True) for t in (x, s, d, w, nz, nw, bias)]
    out = K.StyledConvTCFunction.apply(*dev, 0.2, 2 ** 0.5)
    assert out.shape == ref.shape and (out.detach().cpu().double() - ref.detach()).abs().max() <= 1e-2 * ref.abs().max()
    out.backward(gy.to(DEV))
    for name, a, b in zip(("gx", "gs", "gd", "gw", "gnoise", "gnw", "gbias"), dev, leaves):
        err = (a.grad.cpu().double() - b.grad).abs().max().item()
        assert a.grad.shape == b.grad.shape and err <= 1.2e-2 * b.grad.abs().max().item() + 1e-6, (name, err, b.grad.abs().max().item())
    # frozen parameters (the ReStyle direction): only x, s, d receive gradients
    dev = [t.to(DEV) for t in (x, s, d, w, nz, nw, bias)]
    for t in dev[:3]:
        t.requires_grad_(True)
    K.StyledConvTCFunction.apply(*dev, 0.2, 2 ** 0.5).backward(gy.to(DEV))
    for a, b in zip(dev[:3], leaves[:3]):
        assert (a.grad.cpu().double() - b.grad).abs().max() <= 1.2e-2 * b.grad.abs().max()
    assert all(t.grad is None for t in dev[3:])


@pytest.mark.parametrize("B,Cn,R", [(2, 32, 9), (1, 96, 33), (3, 64, 5)])
def test_polyphase_layout_passes(sg2, B, Cn, R):
    """NCHW <-> the four polyphase planes the transposed convolution writes (no interleaving copies)"""
    K = _K()
    g = torch.Generator().manual_seed(Cn + R)
    x = torch.randn(B, Cn, R, R, generator=g)
    sc = torch.rand(B, Cn, generator=g) + 0.5
    P = (R + 1) // 2
    oth = torch.randn(4, B, P, P, Cn, generator=g).bfloat16()
    planes, red = K.to_planes(x.to(DEV), sc.to(DEV), other=oth.to(DEV))
    xs = (x * sc.view(B, Cn, 1, 1))
    rref = torch.zeros(B, Cn, dtype=torch.float64)
    for s in range(4):
        py, px = s >> 1, s & 1
        want = xs[:, :, py::2, px::2].permute(0, 2, 3, 1)
        got = planes[s].cpu().float()
        assert (got[:, :P - py, :P - px] - want).abs().max() <= 2 ** -8 * want.abs().max()
        assert got[:, P - py:].abs().sum() == 0 and got[:, :, P - px:].abs().sum() == 0       # the padding stays zero
        rref += (x[:, :, py::2, px::2].permute(0, 2, 3, 1).double() * oth[s, :, :P - py, :P - px].double()).sum((1, 2))
    assert (red.cpu().double() - rref).abs().max() <= 1e-5 * rref.abs().max() + 1e-5
    y, red2 = K.planes_to_nchw(planes, sc.to(DEV), torch.float32, other=x.to(DEV))
    yref = torch.zeros(B, Cn, R, R)
    for s in range(4):
        py, px = s >> 1, s & 1
        yref[:, :, py::2, px::2] = planes[s].cpu().float()[:, :P - py, :P - px].permute(0, 3, 1, 2)
    r2 = (yref.double() * x.double()).sum((2, 3))
    assert (y.cpu() - yref * sc.view(B, Cn, 1, 1)).abs().max() <= 1e-6 * yref.abs().max()
    assert (red2.cpu().double() - r2).abs().max() <= 1e-5 * r2.abs().max() + 1e-5


@pytest.mark.parametrize("B,Cn,H,W,dt", [(2, 128, 16, 16, torch.float32), (3, 512, 4, 4, torch.float32), (1, 32, 65, 67, torch.float32),
                                        (2, 64, 32, 32, torch.bfloat16), (1, 13, 130, 70, torch.float32)])
def test_rgb_modconv_function_vs_fp64(sg2, B, Cn, H, W, dt):
    """the modulated 1x1 convolution of ToRGB (model.py:350-355) and every gradient vs fp64 autograd"""
    K = _K()
    g = torch.Generator().manual_seed(Cn + H)
    x = torch.randn(B, Cn, H, W, generator=g).to(dt)
    s = torch.rand(B, Cn, generator=g) + 0.5
    w = torch.randn(3, Cn, generator=g) / Cn ** 0.5
    gy = torch.randn(B, 3, H, W, generator=g).to(dt)
    x64, s64, w64 = (t.double().requires_grad_(True) for t in (x, s, w))
    ref = torch.einsum("kc,bc,bchw->bkhw", w64, s64, x64)
    ref.backward(gy.double())
    xd, sd, wd = (t.to(DEV).requires_grad_(True) for t in (x, s, w))
    y = K.RgbModConvFunction.apply(xd, sd, wd)
    y.backward(gy.to(DEV))
    tol = 2 ** -7 if dt == torch.bfloat16 else 2e-5
    for name, a, b in (("y", y.detach(), ref.detach()), ("gx", xd.grad, x64.grad), ("gs", sd.grad, s64.grad), ("gw", wd.grad, w64.grad)):
        err = (a.cpu().double() - b).abs().max().item()
        assert a.shape == b.shape and err <= tol * b.abs().max().item() + 1e-6, (name, err, b.abs().max().item())


def test_frozen_weight_cache_follows_the_parameter(sg2, oracle):
    """the packed-weight cache of the tensor-core route is keyed on the parameter's version: an in-place update of a
    frozen weight (load_state_dict, optimizer step of an outer loop) must be seen by the next forward"""
    G = sg2.Generator(16, 512, 2).to(DEV).eval()
    G.precision = 'bf16'
    for p in G.parameters():
        p.requires_grad_(False)
    lat = torch.randn(2, G.n_latent, 512, device=DEV)
    a, _ = G([lat.clone().requires_grad_(True)], input_is_latent=True, randomize_noise=False)
    b, _ = G([lat.clone().requires_grad_(True)], input_is_latent=True, randomize_noise=False)
    assert torch.equal(a, b)
    with torch.no_grad():
        G.convs[1].conv.weight.mul_(1.5)
        G.convs[1].conv.weight[0, :8].add_(0.3)
    c, _ = G([lat.clone().requires_grad_(True)], input_is_latent=True, randomize_noise=False)
    G.precision = 'exact'
    e, _ = G([lat.clone().requires_grad_(True)], input_is_latent=True, randomize_noise=False)
    assert not torch.equal(a, c)
    assert (c - e).abs().max() <= 3e-2 * e.abs().max()          # the updated weights, not the cached ones
