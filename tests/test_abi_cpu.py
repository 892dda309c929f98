"""CPU: the C-ABI library loads, exports exactly what include/sg2_b200.h declares, the Python
bindings cover every symbol, and the host-side mirror behaves like the reference's interface
(names, state_dict layout, error behaviour).  No compute calls (no GPU here)."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "sg2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sg2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol(sg2):
    lib = ctypes.CDLL(sg2._lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sg2_b200.h but not exported"


def test_bindings_cover_header(sg2):
    assert sorted(sg2._lib.SIGNATURES) == _header_symbols()
    lib = sg2._lib.load()
    assert lib.sg2_abi_version() == 1


def test_conv_params_struct_layout(sg2):
    # 5 pointers + 6 int32 -> 64 bytes, matching sg2_conv_params in the header
    assert ctypes.sizeof(sg2._lib.ConvParams) == 5 * 8 + 6 * 4


def test_fastdiv_selfcheck(sg2):
    import random
    lib = sg2._lib.load()
    rnd = random.Random(0)
    for _ in range(20000):
        d = rnd.choice([1, 2, 3, 5, 7, 16, 100, 255, 256, 257, 4096, 65536, (1 << 31) - 1, rnd.randint(1, (1 << 31) - 1)])
        n = rnd.choice([0, 1, d - 1, d, d + 1, (1 << 31) - 1, rnd.randint(0, (1 << 31) - 1)])
        assert lib.sg2_selftest_fastdiv(d, max(0, n)) == 0, (d, n)


def test_bad_arguments_are_reported_not_launched(sg2):
    lib = sg2._lib.load()
    before = lib.sg2_launch_count()
    rc = lib.sg2_upfirdn2d(None, None, None, 1, 8, 8, 1, 4, 4, 7, 7, 1, 1, 0, 0, 0, 0, 0, None)
    assert rc == 2 and b"unsupported configuration" in lib.sg2_last_error()     # up=7: reference returns garbage
    rc = lib.sg2_fused_bias_act(None, None, None, None, 16, 1, 1, 3, 0, 0.2, 1.4, 0, None)
    assert rc == 1 and b"null" in lib.sg2_last_error()
    rc = lib.sg2_fused_bias_act(None, None, None, None, 16, 1, 1, 2, 0, 0.2, 1.4, 0, None)
    assert rc in (1, 2)
    rc = lib.sg2_modconv2d_fwd(None, None, None, None, None, 1, 8, 8, 4, 4, 5, 0, 0, None)
    assert rc == 2 and b"kernel size" in lib.sg2_last_error()
    assert lib.sg2_launch_count() == before


def test_ops_reject_cpu_tensors_like_the_reference(sg2):
    # op/fused_bias_act.cpp:7,13-14 and op/upfirdn2d.cpp:8,15-16: "... must be a CUDA tensor"
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        sg2.fused_leaky_relu(torch.zeros(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        sg2.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    G = sg2.Generator(8, 512, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        G([torch.zeros(1, 512)])


def test_state_dict_layout_matches_reference(sg2):
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    for tag, ref in keys.items():
        size, n_mlp, cm = (int(v) for v in tag.split("_"))
        if size > 256:
            continue
        G = sg2.Generator(size, 512, n_mlp, channel_multiplier=cm)
        ours = [(k, list(v.shape)) for k, v in G.state_dict().items()]
        assert ours == [(k, list(s)) for k, s in ref], tag


def test_generator_attributes_and_oracle_weights_load(sg2, oracle):
    G = sg2.Generator(32, 512, 2)
    assert (G.n_latent, G.num_layers, G.size, G.style_dim, G.log_size) == (8, 7, 32, 512, 5)
    sd = oracle.init_state_dict(32, 512, 2)
    missing, unexpected = G.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert G.to_rgbs[0].upsample.pad == (2, 1) and G.convs[0].conv.blur.pad == (1, 1)   # SURVEY 2.2 [probe]
    assert float(G.convs[0].conv.blur.kernel.sum()) == pytest.approx(4.0)
    assert [tuple(n.shape) for n in G.make_noise()] == [(1, 1, 4, 4)] + [(1, 1, r, r) for r in (8, 8, 16, 16, 32, 32)]
    assert isinstance(G.style, torch.nn.Sequential) and len(G.style) == 3


def test_reference_init_statistics(sg2):
    torch.manual_seed(0)
    G = sg2.Generator(16, 512, 2)
    assert float(G.conv1.noise.weight) == 0 and float(G.conv1.activate.bias.abs().sum()) == 0   # fact 6
    assert float(G.to_rgb1.bias.abs().sum()) == 0
    assert float(G.conv1.conv.modulation.bias.mean()) == 1
    assert 80 < float(G.style[1].weight.std()) < 120                                             # randn / 0.01


def test_binding_arity_and_types_match_the_header(sg2):
    """every ctypes signature has as many arguments as the C prototype, pointers where the header has pointers,
    64-bit integers where it has int64_t and floats where it has float (a mismatch corrupts the call silently)"""
    txt = open(os.path.join(ROOT, "include", "sg2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = dict(re.findall(r"\b(sg2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S))
    assert set(protos) == set(sg2._lib.SIGNATURES)
    for name, (restype, argtypes) in sg2._lib.SIGNATURES.items():
        params = [p.strip() for p in protos[name].replace("\n", " ").split(",")]
        if params == ["void"] or params == [""]:
            params = []
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p, t in zip(params, argtypes):
            is_ptr_c = "*" in p or "sg2_stream_t" in p
            is_ptr_py = t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "_type_") and hasattr(t, "contents")
            assert is_ptr_c == bool(is_ptr_py), (name, p, t)
            if not is_ptr_c:
                if "int64_t" in p or "long long" in p:
                    assert t is ctypes.c_int64, (name, p, t)
                elif "float" in p:
                    assert t is ctypes.c_float, (name, p, t)
                elif "uint32_t" in p or "unsigned" in p:
                    assert t in (ctypes.c_uint, ctypes.c_uint32), (name, p, t)
                else:
                    assert t is ctypes.c_int, (name, p, t)
    rets = dict((n, r.strip()) for r, n in re.findall(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_]*(?:\s*\*)?)\s*(sg2_[a-z0-9_]+)\s*\(", txt, flags=re.M))
    for name, (restype, _) in sg2._lib.SIGNATURES.items():
        r = rets[name]
        if "char" in r:
            assert restype is ctypes.c_char_p, (name, r)
        elif r == "void":
            assert restype is None, (name, r)
        elif "int64_t" in r:
            assert restype is ctypes.c_int64, (name, r)
        elif r == "int":
            assert restype is ctypes.c_int, (name, r)
        else:
            raise AssertionError((name, r))
