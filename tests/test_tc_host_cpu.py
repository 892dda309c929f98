"""CPU: host-side logic of the tensor-core route of the differentiable path (no kernel calls): the bf16-operands
switch, the eligibility rule, and that the product functions refuse CPU tensors instead of falling back."""
import importlib
import threading

import pytest
import torch


def _K():
    return importlib.import_module("stylegan-for-facerec_b200.stylegan2.functional")


def test_tc_grad_switch_nests_and_is_thread_local(sg2, monkeypatch):
    K = _K()
    monkeypatch.delenv("SG2_B200_PRECISION", raising=False)
    assert not K.tc_grad_enabled()
    with K.tc_grad(True):
        assert K.tc_grad_enabled()
        with K.tc_grad(None):                    # None leaves the ambient setting alone
            assert K.tc_grad_enabled()
        with K.tc_grad(False):
            assert not K.tc_grad_enabled()
        assert K.tc_grad_enabled()
        seen = []
        t = threading.Thread(target=lambda: seen.append(K.tc_grad_enabled()))
        t.start(); t.join()
        assert seen == [False]                   # another thread (a DataParallel replica) keeps its own setting
    assert not K.tc_grad_enabled()
    monkeypatch.setenv("SG2_B200_PRECISION", "bf16")
    assert K.tc_grad_enabled()                   # environment default for unmodified callers
    with K.tc_grad(False):
        assert not K.tc_grad_enabled()


def test_tc_conv_eligibility(sg2, monkeypatch):
    K = _K()
    monkeypatch.delenv("SG2_B200_PRECISION", raising=False)
    w = torch.zeros(64, 32, 3, 3)
    x = torch.zeros(2, 32, 8, 8)
    assert not K.tc_conv_ok(x, w, 0)                                   # fp32 tensors: bf16 operands must be allowed first
    assert K.tc_conv_ok(x.bfloat16(), w, 0)                            # bfloat16 tensors opt in by themselves
    with K.tc_grad(True):
        assert K.tc_conv_ok(x, w, 0) and K.tc_conv_ok(x, w, 1)
        assert not K.tc_conv_ok(x, w, 2)                               # stride-2 conv (discriminator side): fp32 kernel
        assert not K.tc_conv_ok(x, torch.zeros(64, 32, 1, 1), 0)       # 1x1: ToRGB has its own pass
        assert not K.tc_conv_ok(torch.zeros(2, 32, 8, 6), w, 0)        # square maps only
        assert not K.tc_conv_ok(torch.zeros(2, 32, 2, 2), w, 0)        # r >= 4
        assert not K.tc_conv_ok(torch.zeros(2, 24, 8, 8), torch.zeros(64, 24, 3, 3), 0)    # Cin % 32
        assert not K.tc_conv_ok(x, torch.zeros(48, 32, 3, 3), 0)       # Cout % 32


def test_tensor_core_functions_refuse_cpu_tensors(sg2):
    K = _K()
    x = torch.zeros(1, 32, 8, 8)
    for fn in (lambda: K.to_nhwc(x), lambda: K.tc_conv3x3(x, torch.zeros(32, 32, 3, 3)),
               lambda: K.RgbModConvFunction.apply(x, torch.ones(1, 32), torch.zeros(3, 32))):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_generators_expose_the_precision_switch(sg2, monkeypatch):
    monkeypatch.delenv("SG2_B200_PRECISION", raising=False)
    G = sg2.Generator(8, 32, 1)
    assert G.precision == "auto"
    A = sg2.stylegan2_ada.Generator(32, 32, 1, 8, 3)
    assert A.precision == "auto" and A.synthesis.precision == "auto"
    A.precision = "bf16"
    assert A.synthesis.precision == "bf16" and "precision" not in A.state_dict()
    monkeypatch.setenv("SG2_B200_PRECISION", "bf16")
    assert sg2.Generator(8, 32, 1).precision == "bf16" and sg2.stylegan2_ada.Generator(32, 32, 1, 8, 3).precision == "bf16"


def test_frozen_pack_cache_keys_on_the_parameter_version(sg2, monkeypatch):
    """ModulatedConv2d._tc_weights: cached while the weight is frozen and unchanged, rebuilt after an in-place update,
    bypassed (and differentiable) while the weight trains.  The pack kernel is replaced by a CPU stand-in."""
    T = importlib.import_module("stylegan-for-facerec_b200.stylegan2.tc_route")
    calls = []
    monkeypatch.setattr(T, "_tc_pack", lambda w: (calls.append(tuple(w.shape)), w.detach().clone())[1])
    conv = sg2.ModulatedConv2d(32, 64, 3, 16)
    conv.weight.requires_grad_(False)
    w4, wp, wp_adj, wsq = conv._tc_weights()
    assert len(calls) == 2 and w4.shape == (64, 32, 3, 3) and wp_adj.shape == (32, 64, 3, 3) and wsq.shape == (64, 32)
    assert torch.allclose(w4, conv.weight[0] * conv.scale) and torch.allclose(wsq, w4.pow(2).sum([2, 3]))
    assert torch.equal(wp_adj, w4.flip([2, 3]).transpose(0, 1))          # adjoint of the 'same' conv: taps flipped, roles swapped
    again = conv._tc_weights()
    assert len(calls) == 2 and all(a is b for a, b in zip(again, (w4, wp, wp_adj, wsq)))      # served from the cache
    with torch.no_grad():
        conv.weight.mul_(2.0)                                            # in-place update bumps the version counter
    w4b, _, _, wsqb = conv._tc_weights()
    assert len(calls) == 4 and torch.allclose(w4b, 2 * w4) and torch.allclose(wsqb, 4 * wsq)
    conv.weight.requires_grad_(True)                                     # training the decoder: no cache, gradients flow
    w4c, wpc, wpac, wsqc = conv._tc_weights()
    assert len(calls) == 4 and wpc is None and wpac is None and w4c.requires_grad and wsqc.requires_grad
    with torch.no_grad():                                                # ... unless autograd is off altogether
        assert conv._tc_weights()[1] is not None
    up = sg2.ModulatedConv2d(32, 64, 3, 16, upsample=True)
    up.weight.requires_grad_(False)
    w4u, _, wp_adj_u, _ = up._tc_weights()
    assert torch.equal(wp_adj_u, w4u.transpose(0, 1))                    # adjoint of the transposed conv: same taps
