"""GPU parity tests of the op API (through the C ABI) against the oracle and the golden vectors."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tol(dtype):
    return {torch.float32: 0.0, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dtype]


# ---- fused_leaky_relu -------------------------------------------------------------------------------
def test_lrelu_golden_bit_exact(sg2, oracle, golden, cases):
    for name, shape in cases.LRELU_CASES:
        x = oracle.named_randn("lrelu:" + name, shape, 3)
        b = oracle.named_randn("lrelu_b:" + name, (shape[1],), 3)
        y = sg2.fused_leaky_relu(x.to(DEV), b.to(DEV)).cpu()
        assert torch.equal(y, torch.from_numpy(golden["ops"]["lrelu/" + name])), name   # fp32: bit-exact


@pytest.mark.parametrize("shape", [(4, 512), (3, 7), (2, 512, 4, 4), (2, 64, 33, 31), (1, 3, 256, 256),
                                   (5, 17, 9), (2, 32, 128, 128), (1, 1, 1, 1), (0, 8, 4, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_lrelu_vs_oracle(sg2, oracle, shape, dtype):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g).to(dtype)
    b = torch.randn(shape[1], generator=g).to(dtype)
    y = sg2.fused_leaky_relu(x.to(DEV), b.to(DEV), 0.1, 1.7)
    assert y.dtype == dtype and y.shape == x.shape
    ref = oracle.fused_leaky_relu(x.float(), b.float(), 0.1, 1.7)
    if dtype == torch.float32:
        assert torch.equal(y.cpu(), ref)
    else:
        assert torch.equal(y.cpu(), ref.to(dtype))      # fp32 math then one rounding -> same bits


def test_lrelu_all_modes_vs_oracle(sg2, oracle):
    from importlib import import_module
    fa = import_module("stylegan-for-facerec_b200.stylegan2.op.fused_act")
    g = torch.Generator().manual_seed(2)
    x, ref, b = torch.randn(3, 10, 6, 6, generator=g), torch.randn(3, 10, 6, 6, generator=g), torch.randn(10, generator=g)
    for act in (1, 3):
        for grad in (0, 1, 2):
            y = fa.bias_act(x.to(DEV), b.to(DEV), ref.to(DEV), act, grad, 0.2, 1.5).cpu()
            assert torch.equal(y, oracle.bias_act(x, b, ref, act, grad, 0.2, 1.5)), (act, grad)
    y = fa.bias_act(x.to(DEV), None, None, 3, 0, 0.2, 1.0).cpu()
    assert torch.equal(y, oracle.bias_act(x, None, None, 3, 0, 0.2, 1.0))


def test_lrelu_noncontiguous_and_unaligned(sg2, oracle):
    x = torch.randn(4, 6, 10, 10)
    b = torch.randn(6)
    xt = x.to(DEV).permute(0, 1, 3, 2)             # non-contiguous view
    assert torch.equal(sg2.fused_leaky_relu(xt, b.to(DEV)).cpu(), oracle.fused_leaky_relu(x.permute(0, 1, 3, 2), b))
    big = torch.randn(4 * 6 * 9 * 9 + 1, device=DEV)
    xo = big[1:].view(4, 6, 9, 9)                   # 4-byte aligned only -> scalar path
    assert torch.equal(sg2.fused_leaky_relu(xo, b.to(DEV)).cpu(), oracle.fused_leaky_relu(xo.cpu(), b))


def test_lrelu_backward_and_double_backward(sg2, oracle, golden, cases):
    for name, shape in cases.LRELU_CASES:
        x = oracle.named_randn("lrelu:" + name, shape, 3).to(DEV).requires_grad_(True)
        b = oracle.named_randn("lrelu_b:" + name, (shape[1],), 3).to(DEV).requires_grad_(True)
        gy = oracle.named_randn("lrelu_gy:" + name, shape, 3).to(DEV)
        y = sg2.fused_leaky_relu(x, b)
        gx, gb = torch.autograd.grad(y, [x, b], gy, create_graph=True)
        np.testing.assert_allclose(gx.detach().cpu().numpy(), golden["ops"]["lrelu_gx/" + name], rtol=0, atol=1e-6)
        np.testing.assert_allclose(gb.detach().cpu().numpy(), golden["ops"]["lrelu_gb/" + name], rtol=1e-5, atol=1e-5)
        # second order: d/dgy of (gx . v) = act'(out) * v * scale  (op/fused_act.py:41-47)
        gy2 = gy.clone().requires_grad_(True)
        gx2, _ = torch.autograd.grad(sg2.fused_leaky_relu(x, b), [x, b], gy2, create_graph=True)
        v = torch.randn_like(gx2)
        ggy, = torch.autograd.grad(gx2, gy2, v)
        mask = (y.detach() > 0).float()
        expect = v * (mask + (1 - mask) * 0.2) * 2 ** 0.5
        np.testing.assert_allclose(ggy.cpu().numpy(), expect.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_lrelu_module(sg2):
    m = sg2.FusedLeakyReLU(8).to(DEV)
    assert list(m.state_dict()) == ["bias"] and m.negative_slope == 0.2 and m.scale == 2 ** 0.5
    x = torch.randn(2, 8, 5, 5, device=DEV)
    assert torch.equal(m(x), torch.nn.functional.leaky_relu(x, 0.2) * (2 ** 0.5))


@pytest.mark.parametrize("n", [64 * 128 * 256 * 256])      # BASELINE config 2's largest activation
def test_lrelu_full_size_properties(sg2, n):
    """size-independent properties at full size: positive homogeneity and the exact 0.2 ratio."""
    x = torch.randn(64, 128, 256, 256, device=DEV, dtype=torch.bfloat16)
    zero = torch.zeros(128, device=DEV, dtype=torch.bfloat16)
    y = sg2.fused_leaky_relu(x, zero, 0.25, 2.0)       # powers of two: exact in bf16
    assert torch.equal(y, torch.where(x > 0, x * 2, x * 0.5))
    del y
    torch.cuda.empty_cache()


# ---- upfirdn2d ----------------------------------------------------------------------------------------
def test_upfirdn2d_golden(sg2, oracle, golden, cases):
    for name, shape, kspec, up, down, pad in cases.UPFIRDN_CASES:
        x = oracle.named_randn("upfirdn:" + name, shape, 3)
        y = sg2.upfirdn2d(x.to(DEV), cases.fir(kspec).to(DEV), up, down, pad).cpu()
        ref = torch.from_numpy(golden["ops"]["upfirdn2d/" + name])
        assert y.shape == ref.shape, name
        np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=3e-6, err_msg=name)


UFD_SWEEP = [
    # shape, k, up, down, pad
    ((2, 8, 257, 257), 4, 1, 1, (1, 1)), ((2, 3, 128, 128), 4, 2, 1, (2, 1)), ((2, 8, 256, 256), 4, 1, 2, (1, 1)),
    ((3, 5, 4, 4), 4, 1, 1, (1, 1)), ((3, 5, 4, 4), 4, 2, 1, (2, 1)), ((1, 2, 5, 5), 4, 1, 2, (1, 1)),
    ((1, 4, 33, 65), 3, 1, 1, (1, 1)), ((1, 4, 33, 65), 2, 2, 1, (1, 0)), ((1, 3, 40, 24), 4, 2, 1, (1, 2)),
    ((1, 3, 40, 24), 4, 2, 1, (3, 0)), ((1, 2, 31, 31), 5, 1, 1, (2, 2)), ((1, 2, 20, 20), 6, 3, 1, (3, 2)),
    ((1, 2, 20, 20), 4, 2, 2, (1, 1)), ((1, 2, 30, 30), 7, 1, 3, (3, 3)), ((2, 2, 16, 16), 4, 1, 1, (-1, -1)),
    ((1, 1, 9, 9), 1, 1, 1, (0, 0)), ((1, 2, 12, 12), 12, 1, 1, (6, 5)), ((1, 2, 16, 16), 4, 4, 1, (3, 0)),
    # row-streaming kernel: several strips / bands / groups per warp, odd pitches (1-D TMA view), every pad parity
    ((1, 2, 300, 300), 4, 1, 1, (1, 1)), ((3, 2, 131, 259), 4, 1, 1, (2, 2)), ((2, 3, 40, 20), 4, 1, 1, (1, 1)),
    ((1, 2, 270, 136), 3, 1, 1, (0, 2)), ((2, 2, 150, 72), 4, 2, 1, (2, 1)), ((1, 3, 150, 72), 4, 2, 1, (1, 2)),
    ((1, 2, 67, 67), 4, 2, 1, (3, 0)), ((1, 2, 130, 30), 2, 2, 1, (0, 1)), ((2, 2, 260, 264), 4, 1, 2, (1, 1)),
    ((1, 3, 129, 67), 4, 1, 2, (2, 2)), ((1, 2, 64, 520), 4, 1, 2, (0, 1)), ((5, 7, 17, 17), 4, 1, 1, (1, 1)),
    ((4, 9, 32, 32), 4, 2, 1, (2, 1)), ((4, 9, 64, 64), 4, 1, 2, (1, 1)), ((1, 1, 1030, 1030), 4, 1, 1, (1, 1)),
    # fp32 down-sampling with 16-byte rows (vector loads): every offset of the staged line, several strips, a plane narrower than a strip
    ((2, 3, 128, 136), 4, 1, 2, (2, 2)), ((1, 2, 96, 72), 4, 1, 2, (3, 1)), ((3, 2, 40, 24), 3, 1, 2, (0, 0)), ((1, 2, 70, 300), 4, 1, 2, (1, 1)),
    ((2, 2, 33, 16), 2, 1, 2, (0, 0)),
]


@pytest.mark.parametrize("shape,k,up,down,pad", UFD_SWEEP)
def test_upfirdn2d_sweep_vs_oracle(sg2, oracle, shape, k, up, down, pad):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=g)
    taps = torch.randn(k, k, generator=g)             # asymmetric: the flip matters
    y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), up, down, pad).cpu()
    ref = oracle.upfirdn2d(x.double(), taps.double(), up, down, pad).float()
    assert y.shape == ref.shape
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=2e-5 * k)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,up,down,pad", [((2, 3, 257, 257), 1, 1, (1, 1)), ((2, 3, 128, 128), 2, 1, (2, 1)),
                                               ((2, 3, 256, 256), 1, 2, (1, 1)), ((1, 2, 131, 70), 1, 1, (2, 2)),
                                               ((1, 2, 37, 45), 2, 1, (1, 2)), ((1, 2, 77, 141), 1, 2, (2, 2))])
def test_upfirdn2d_streaming_low_precision(sg2, oracle, dtype, shape, up, down, pad):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g).to(dtype)
    taps = torch.randn(4, 4, generator=g)
    y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), up, down, pad)
    assert y.dtype == dtype
    ref = oracle.upfirdn2d(x.double(), taps.double(), up, down, pad)
    tol = _tol(dtype) * float(ref.abs().max())
    np.testing.assert_allclose(y.float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=tol)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,pad", [((2, 3, 40, 72), (2, 1)), ((1, 2, 33, 256), (2, 2)), ((3, 5, 20, 8), (2, 1)), ((1, 2, 65, 300), (2, 1)),
                                       ((2, 2, 128, 128), (2, 3)), ((1, 3, 17, 16), (2, 0))])
def test_upfirdn2d_up2_vector_rows_low_precision(sg2, oracle, dtype, shape, pad):
    """up-sampling in 2-byte storage with pad_x0 = 2 on planes whose rows are multiples of 8 bytes: rows fetched as 8-byte vectors,
    6 of them in flight, the staged line starting 3 elements before the strip.  The input is the very end of an allocation that
    is NaN everywhere else: nothing outside the tensor may be read; both row parities of the padding."""
    g = torch.Generator().manual_seed(21)
    n = int(np.prod(shape))
    big = torch.full((n + 4096 + 8,), float("nan"), device=DEV, dtype=dtype)
    off = big.numel() - n
    off -= off % 8
    x = big[off:off + n].view(shape)
    x.copy_(torch.randn(shape, generator=g).to(dtype))
    taps = torch.randn(4, 4, generator=g)
    y = sg2.upfirdn2d(x, taps.to(DEV), 2, 1, pad)
    ref = oracle.upfirdn2d(x.cpu().double(), taps.double(), 2, 1, pad)
    assert y.dtype == dtype and y.shape == ref.shape and torch.isfinite(y.float()).all()
    np.testing.assert_allclose(y.float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=_tol(dtype) * float(ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,down,pad", [((2, 3, 257, 257), 1, (1, 1)), ((1, 2, 300, 131), 1, (2, 2)), ((3, 4, 17, 17), 1, (1, 1)),
                                            ((2, 3, 256, 256), 2, (1, 1)), ((1, 2, 129, 67), 2, (2, 2)), ((2, 2, 24, 20), 2, (1, 1))])
def test_upfirdn2d_separable_taps(sg2, oracle, dtype, shape, down, pad):
    """outer-product taps (what make_kernel builds, model.py:39-48): in 2-byte storage the packed kernel (upfirdn2d_pk.cu) runs the
    blur as a horizontal pass + vertical accumulation, decided on the device from the taps; asymmetric factors so that a flipped
    or transposed factor would show; 3-tap and 4-tap factors; and a rank-2 kernel close to separable stays on the 16-tap path
    with the same result"""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(shape, generator=g).to(dtype)
    for ky, kx in ((torch.tensor([1.0, -2.0, 3.5, 0.25]), torch.tensor([0.5, 4.0, -1.0, 2.0])),
                   (torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8, torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8),
                   (torch.tensor([0.3, -1.0, 2.0]), torch.tensor([1.5, 0.0, -0.7]))):
        taps = torch.outer(ky, kx)
        y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), 1, down, pad)
        ref = oracle.upfirdn2d(x.double(), taps.double(), 1, down, pad)
        assert y.dtype == dtype and y.shape == ref.shape
        tol = (_tol(dtype) if dtype != torch.float32 else 2e-6) * float(ref.abs().max())
        np.testing.assert_allclose(y.float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=tol)
        bumped = taps.clone()
        bumped[1, 2] += 1e-3 * float(taps.abs().max())      # no longer an outer product
        y2 = sg2.upfirdn2d(x.to(DEV), bumped.to(DEV), 1, down, pad)
        ref2 = oracle.upfirdn2d(x.double(), bumped.double(), 1, down, pad)
        np.testing.assert_allclose(y2.float().cpu().numpy(), ref2.float().numpy(), rtol=0, atol=tol)


PK_SHAPES = [
    # planes, h, w: every alignment of the row segment (odd pitches), 8 / 16 / 32 lanes per strip, several strips and bands
    (3, 40, 65), (2, 65, 66), (3, 17, 67), (2, 130, 69), (3, 40, 71), (5, 129, 129), (2, 65, 131), (3, 40, 200), (2, 257, 257),
    (3, 65, 258), (2, 40, 261), (1, 130, 300), (1, 513, 513), (2, 64, 520), (1, 70, 1025), (4, 256, 256), (37, 64, 64), (13, 33, 129),
    (9500, 33, 33), (9473, 17, 40),         # 4 lanes per strip: only with thousands of planes (8 planes per warp)
    (19000, 17, 17), (18950, 9, 20),        # 2 lanes per strip (16 planes per warp)
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("planes,h,w", PK_SHAPES)
def test_upfirdn2d_packed_kernel_bit_identical_to_streaming(sg2, oracle, monkeypatch, dtype, planes, h, w):
    """upfirdn2d_pk.cu (2-byte storage, blur and down-2: packed fp32 FMAs, 16-byte cp.async rows at any pitch) applies the taps in
    the order of the row-streaming kernel (= the reference kernel's, upfirdn2d_kernel.cu:104-117) with fused multiply-adds: with
    taps that are not an outer product the two kernels must agree BIT FOR BIT, whatever the alignment mode, the grouping of
    planes / bands inside a warp, or the position of the tensor in its allocation (over-reads would pick up the NaNs around it);
    and both agree with the fp64 oracle to the storage precision"""
    g = torch.Generator().manual_seed(h * 1000 + w)
    for down, pad, k in ((1, (1, 1), 4), (1, (2, 2), 4), (1, (0, 2), 3), (1, (3, 0), 4), (2, (1, 1), 4), (2, (2, 2), 4), (2, (0, 1), 4)):
        taps = torch.randn(k, k, generator=g)
        n = planes * h * w
        big = torch.full((n + 4096 + 8,), float("nan"), device=DEV, dtype=dtype)
        off = big.numel() - n
        off -= off % 8                                   # 16-byte aligned, at most 7 NaN elements behind the tensor
        x = big[off:off + n].view(1, planes, h, w)
        x.copy_(torch.randn(1, planes, h, w, generator=g).to(dtype))
        outs = {}
        for name, env in (("packed", {"SG2_UPFIRDN_PK": "1"}), ("packed, per-row alignment switch", {"SG2_UPFIRDN_PK": "1", "SG2_UPFIRDN_PK_QS": "-1"}),
                          ("packed, planes 8 apart", {"SG2_UPFIRDN_PK": "1", "SG2_UPFIRDN_PK_BYPLANES": "1"}), ("streaming", {"SG2_UPFIRDN_PK": "0"})):
            with monkeypatch.context() as m:
                for key, val in env.items():
                    m.setenv(key, val)
                outs[name] = sg2.upfirdn2d(x, taps.to(DEV), 1, down, pad)
        ref = oracle.upfirdn2d(x.cpu().double(), taps.double(), 1, down, pad)
        for name, y in outs.items():
            assert y.dtype == dtype and y.shape == ref.shape
            assert torch.equal(y.view(torch.int16), outs["streaming"].view(torch.int16)), (name, down, pad, k)
        tol = _tol(dtype) * float(ref.abs().max())
        np.testing.assert_allclose(outs["packed"].float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=tol)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_upfirdn2d_packed_kernel_separable_blur(sg2, oracle, monkeypatch, dtype):
    """make_kernel taps (an exact outer product) on the model's blur geometry: the separable blur of the packed kernel against its
    own 16-tap path -- at most one unit in the last place of the storage type apart -- and against the fp64 oracle"""
    g = torch.Generator().manual_seed(9)
    taps = (sg2.make_kernel([1, 3, 3, 1]) * 4).to(DEV)
    for shape, pad in (((2, 16, 257, 257), (1, 1)), ((3, 7, 129, 129), (1, 1)), ((2, 5, 65, 65), (1, 1)), ((1, 3, 256, 256), (2, 2))):
        x = torch.randn(shape, generator=g).to(dtype).to(DEV)
        y = sg2.upfirdn2d(x, taps, 1, 1, pad)
        with monkeypatch.context() as m:
            m.setenv("SG2_UPFIRDN_PK_SEP", "0")
            y16 = sg2.upfirdn2d(x, taps, 1, 1, pad)
        ref = oracle.upfirdn2d(x.cpu().double(), taps.cpu().double(), 1, 1, pad)
        eps = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
        lim = eps * torch.maximum(y.float().abs(), y16.float().abs()) + 4e-6 * float(ref.abs().max())
        assert bool(((y.float() - y16.float()).abs() <= lim).all())
        np.testing.assert_allclose(y.float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=_tol(dtype) * float(ref.abs().max()))


UFD_PLANES = [
    # (planes as N, C), h, w, k, up, down, pad -- many small planes: the batched-planes kernel (upfirdn2d_planes.cu)
    ((8, 16), 4, 4, 4, 1, 1, (1, 1)), ((8, 16), 4, 4, 4, 2, 1, (2, 1)), ((6, 7), 5, 5, 4, 1, 1, (2, 2)), ((8, 8), 8, 8, 4, 1, 2, (1, 1)),
    ((3, 50), 9, 9, 4, 1, 1, (1, 1)), ((4, 40), 8, 8, 4, 2, 1, (2, 1)), ((4, 33), 16, 16, 4, 2, 1, (1, 2)), ((2, 64), 17, 17, 4, 1, 1, (1, 1)),
    ((2, 64), 16, 16, 4, 1, 2, (1, 1)), ((3, 21), 33, 33, 4, 1, 1, (1, 1)), ((3, 21), 32, 32, 4, 2, 1, (2, 1)), ((3, 21), 32, 32, 4, 1, 2, (1, 1)),
    ((2, 19), 40, 37, 3, 1, 1, (1, 1)), ((2, 19), 31, 40, 2, 2, 1, (1, 0)), ((2, 19), 24, 30, 4, 1, 2, (2, 2)), ((2, 19), 13, 6, 4, 2, 1, (3, 0)),
    ((2, 19), 12, 12, 4, 1, 1, (-1, -1)), ((2, 19), 10, 14, 4, 1, 1, (3, 0)), ((2, 19), 10, 14, 4, 2, 1, (3, 3)), ((1, 16), 7, 7, 1, 1, 1, (0, 0)),
    ((5, 7), 6, 6, 4, 1, 1, (4, 4)), ((5, 7), 6, 6, 4, 1, 1, (5, 5)),      # pad 5 leaves the bordered tile: falls through to the other kernels
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("nc,h,w,k,up,down,pad", UFD_PLANES)
def test_upfirdn2d_small_planes(sg2, oracle, dtype, nc, h, w, k, up, down, pad):
    """SG2_UPFIRDN_PLANES=2 sends every shape the batched-planes kernel can express through it (the default keeps it to the sizes
    where it measured faster), =0 switches it off: both against the oracle, and against each other"""
    import os
    g = torch.Generator().manual_seed(17)
    x = torch.randn(nc + (h, w), generator=g).to(dtype)
    taps = torch.randn(k, k, generator=g)             # asymmetric: flips / transposes would show
    ref = oracle.upfirdn2d(x.double(), taps.double(), up, down, pad)
    tol = (_tol(dtype) if dtype != torch.float32 else 4e-6) * max(float(ref.abs().max()), 1.0)
    ys = {}
    for mode in ("2", "0", None):
        if mode is None:
            os.environ.pop("SG2_UPFIRDN_PLANES", None)
        else:
            os.environ["SG2_UPFIRDN_PLANES"] = mode
        try:
            y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), up, down, pad)
        finally:
            os.environ.pop("SG2_UPFIRDN_PLANES", None)
        assert y.dtype == dtype and y.shape == ref.shape
        np.testing.assert_allclose(y.float().cpu().numpy(), ref.float().numpy(), rtol=0, atol=tol, err_msg=f"mode {mode}")
        ys[mode] = y.float().cpu().numpy()
    if dtype == torch.float32:                         # same results as the kernels it replaces on these shapes
        np.testing.assert_allclose(ys["2"], ys["0"], rtol=0, atol=4e-6 * max(float(ref.abs().max()), 1.0))


def test_upfirdn2d_streaming_tensor_edges(sg2, oracle):
    # planes narrower than the staged line: the unpredicated row fetch must not leave the tensor at either end
    # (the input is the tail of an allocation, so an over-read would fault or pick up NaNs)
    for shape, up, down, pad in (((3, 5, 8, 8), 2, 1, (2, 1)), ((2, 3, 9, 9), 1, 1, (2, 2)), ((1, 2, 17, 9), 1, 1, (1, 1)),
                                 ((2, 2, 24, 20), 1, 2, (1, 1))):
        n = int(np.prod(shape))
        big = torch.full((n + 4096,), float("nan"), device=DEV)
        x = big[2048:2048 + n].view(shape)
        x.copy_(torch.randn(shape))
        taps = torch.randn(4, 4)
        y = sg2.upfirdn2d(x, taps.to(DEV), up, down, pad).cpu()
        ref = oracle.upfirdn2d(x.cpu().double(), taps.double(), up, down, pad).float()
        assert torch.isfinite(y).all()
        np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=1e-4)


def test_upfirdn2d_unaligned_base_falls_back(sg2, oracle):
    # a storage offset of one element: not 16-byte aligned -> the TMA path declines, the tiled kernel answers
    big = torch.randn(2 * 3 * 100 * 100 + 1, device=DEV)
    x = big[1:].view(2, 3, 100, 100)
    taps = torch.randn(4, 4)
    y = sg2.upfirdn2d(x, taps.to(DEV), 1, 1, (1, 1)).cpu()
    ref = oracle.upfirdn2d(x.cpu().double(), taps.double(), 1, 1, (1, 1)).float()
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_upfirdn2d_low_precision_accumulates_in_fp32(sg2, oracle, dtype):
    x = torch.randn(2, 4, 64, 64).to(dtype)
    taps = sg2.make_kernel([1, 3, 3, 1]) * 4
    for up, down, pad in ((1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (1, 1))):
        y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), up, down, pad)
        assert y.dtype == dtype
        ref = oracle.upfirdn2d(x.float(), taps, up, down, pad)
        assert torch.equal(y.cpu(), ref.to(dtype)) or (y.cpu().float() - ref).abs().max() <= _tol(dtype) * ref.abs().max()


def test_upfirdn2d_rectangular_taps_and_minor_dim(sg2, oracle):
    from importlib import import_module
    um = import_module("stylegan-for-facerec_b200.stylegan2.op.upfirdn2d")
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 10, 12, 5, generator=g)                       # [major, h, w, minor]
    taps = torch.randn(3, 4, generator=g)
    y = um.upfirdn2d_raw(x.to(DEV), taps.to(DEV), 2, 1, 1, 2, 1, 2, 0, 1).cpu()
    xr = x.permute(0, 3, 1, 2).reshape(1, 15, 10, 12)
    ref = oracle.upfirdn2d_xy(xr, taps, 2, 1, 1, 2, 1, 2, 0, 1)
    ref = ref.reshape(3, 5, ref.shape[2], ref.shape[3]).permute(0, 2, 3, 1)
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=0, atol=1e-5)


def test_upfirdn2d_unsupported_config_raises(sg2):
    x = torch.randn(1, 1, 8, 8, device=DEV)
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        sg2.upfirdn2d(x, torch.ones(4, 4, device=DEV), up=5)          # reference: silent garbage
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        sg2.upfirdn2d(x, torch.ones(17, 17, device=DEV), pad=(8, 8))
    with pytest.raises(RuntimeError, match="empty output"):
        sg2.upfirdn2d(x, torch.ones(4, 4, device=DEV), pad=(-4, -4))


def test_upfirdn2d_backward_is_the_adjoint(sg2, oracle):
    g = torch.Generator().manual_seed(5)
    for shape, k, up, down, pad in [((2, 3, 17, 17), 4, 1, 1, (1, 1)), ((2, 3, 8, 8), 4, 2, 1, (2, 1)),
                                    ((1, 2, 16, 16), 4, 1, 2, (1, 1)), ((1, 2, 9, 7), 3, 2, 1, (1, 1))]:
        x = torch.randn(shape, generator=g)
        taps = torch.randn(k, k, generator=g)
        xd = x.to(DEV).requires_grad_(True)
        y = sg2.upfirdn2d(xd, taps.to(DEV), up, down, pad)
        gy = torch.randn(y.shape, generator=g)
        gx, = torch.autograd.grad(y, xd, gy.to(DEV), create_graph=True)
        xo = x.double().requires_grad_(True)
        go, = torch.autograd.grad(oracle.upfirdn2d(xo, taps.double(), up, down, pad), xo, gy.double())
        np.testing.assert_allclose(gx.detach().cpu().numpy(), go.float().numpy(), rtol=0, atol=1e-4)
        # double backward: linear op -> d(gx.v)/dgy = upfirdn2d(v)
        gyd = gy.to(DEV).requires_grad_(True)
        gx2, = torch.autograd.grad(sg2.upfirdn2d(xd, taps.to(DEV), up, down, pad), xd, gyd, create_graph=True)
        v = torch.randn(shape, generator=g)
        ggy, = torch.autograd.grad(gx2, gyd, v.to(DEV))
        np.testing.assert_allclose(ggy.cpu().numpy(), oracle.upfirdn2d(v.double(), taps.double(), up, down, pad).float().numpy(),
                                   rtol=0, atol=1e-4)


def test_upfirdn2d_full_size_properties(sg2):
    """BASELINE config 2's largest blur: [64*128, 257, 257] -> 256^2, bf16.  Properties that hold
    at any size: DC gain (taps sum to 4 -> constant in, 4*constant out away from the border),
    linearity, and up2 of a constant image reproduces the constant in the interior."""
    taps = (sg2.make_kernel([1, 3, 3, 1]) * 4).to(DEV)
    x = torch.full((64, 128, 257, 257), 0.5, device=DEV, dtype=torch.bfloat16)
    y = sg2.upfirdn2d(x, taps, pad=(1, 1))
    assert y.shape == (64, 128, 256, 256)
    assert torch.all(y[:, :, 1:-1, 1:-1] == 2.0)
    del x, y
    torch.cuda.empty_cache()
    a = torch.randn(8, 16, 128, 128, device=DEV)
    b = torch.randn(8, 16, 128, 128, device=DEV)
    ya, yb, yab = (sg2.upfirdn2d(t, taps, up=2, pad=(2, 1)) for t in (a, b, a + 2 * b))
    assert (yab - (ya + 2 * yb)).abs().max() < 2e-5
    c = sg2.upfirdn2d(torch.ones(1, 1, 64, 64, device=DEV), taps, up=2, pad=(2, 1))
    assert torch.allclose(c[:, :, 2:-2, 2:-2], torch.ones(1, 1, 124, 124, device=DEV))
