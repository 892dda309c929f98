"""CPU: host-side logic of the engines (no kernel launches): the launch planner's choices per layer, the training-mode and
ADA plans, and the lifetime rules of the engine objects attached to a Generator (deepcopy / pickle / DataParallel replicas,
precision validation, explicit cache invalidation)."""
import copy
import ctypes as C
import importlib
import io

import pytest
import torch


def _plan(sg2, G, batch, ada=False, train=False):
    """build a launch plan without a GPU (sg2_synth_create does not touch the device) -> (describe text, workspace bytes)"""
    E = importlib.import_module("stylegan-for-facerec_b200.engine")
    L = sg2._lib
    lib = L.load()
    if ada:
        eng = E.AdaSynthesisEngine.__new__(E.AdaSynthesisEngine)
        eng.G = E._AdaView(G.synthesis)
    else:
        eng = E.SynthesisEngine.__new__(E.SynthesisEngine)
        eng.G = G
    eng.lib = lib
    rows, const, taps, _ = eng._layer_table()
    arr = (L.ConvParams * len(rows))(*rows)
    plan = C.c_void_p()
    rc = eng._create_fn()(C.byref(plan), eng.G.size, eng.G.style_dim, batch, arr, len(rows), const,
                          taps.numpy().ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0, lib.sg2_last_error()
    ws0 = lib.sg2_synth_workspace_bytes(plan)
    if train:
        rc = lib.sg2_synth_enable_training(plan)
        if rc != 0:
            msg = lib.sg2_last_error().decode()
            lib.sg2_synth_destroy(plan)
            return msg, ws0, None
    buf = C.create_string_buffer(1 << 16)
    lib.sg2_synth_describe(plan, buf, len(buf))
    ws1 = lib.sg2_synth_workspace_bytes(plan)
    lib.sg2_synth_destroy(plan)
    return buf.value.decode(), ws0, ws1


def test_planner_choices_256_and_1024(sg2):
    d256, _, _ = _plan(sg2, sg2.Generator(256, 512, 2), 64)
    rows = [l.split() for l in d256.strip().splitlines()]
    kinds = [r[1] for r in rows]
    assert kinds.count("gemm") == 13 and kinds.count("upfir") == 6 and kinds.count("rgb_combine") == 7
    assert not any("_upfused" in r[2] or "_dxs" in r[2] for r in rows)          # nothing narrow enough at 256^2
    d1024, _, _ = _plan(sg2, sg2.Generator(1024, 512, 2), 32)
    names = [l.split()[2] for l in d1024.strip().splitlines()]
    # the two narrow octaves: blur folded into the up-conv (no FIR row follows), the last conv on the dx-stacked kernel
    assert "L20_512x512_128->64_upfused" in names and "L23_1024x1024_64->32_upfused" in names
    assert "L24_1024x1024_32->32_dxs" in names and "L21_512x512_64->64" in names
    assert sum(l.split()[1] == "upfir" for l in d1024.strip().splitlines()) == 6
    # algorithmic FLOPs are the reference's formulation whatever the plan (148.52 GFLOP per image)
    fl = sum(float(dict(f.split("=") for f in l.split()[3:])["flops"]) for l in d1024.strip().splitlines() if l.split()[1] == "gemm")
    assert abs(fl / 1e9 - 148.2) < 0.5


def test_training_plan_grows_the_workspace_and_ada_plan(sg2):
    G = sg2.Generator(64, 512, 2)
    _, ws0, ws1 = _plan(sg2, G, 4, train=True)
    assert ws1 > ws0                                                      # kept activations + adjoint weight packs
    ada = importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator")
    GA = ada.Generator(512, 512, 2, 64, 3)
    d, _, _ = _plan(sg2, GA, 4, ada=True)
    kinds = [l.split()[1] for l in d.strip().splitlines()]
    assert kinds.count("smoothup") == 4 and kinds.count("upfir") == 0 and kinds.count("gemm") == 9
    msg, _, ws = _plan(sg2, GA, 4, ada=True, train=True)
    assert ws is None and "stylegan2_ada" in msg                          # no backward walk for the ADA plan: explicit error


def test_engine_objects_do_not_travel_with_the_module(sg2):
    G = sg2.Generator(16, 512, 2)

    class FakeEngine:                       # what SynthesisEngine holds: a ctypes handle (unpicklable) and a device
        def __init__(self):
            self.plan, self.device = C.c_void_p(5), torch.device("cpu")

    G._engine = FakeEngine()
    G.__dict__["_train_engine"] = FakeEngine()
    G2 = copy.deepcopy(G)
    assert G2._engine is None and "_train_engine" not in G2.__dict__ and G._engine is not None
    torch.save(G, io.BytesIO())                                           # whole-module save works after a bf16 forward
    rep = G._replicate_for_data_parallel()
    assert rep._engine is None and rep.__dict__.get("_transient_engine") and "_train_engine" not in rep.__dict__
    G.invalidate_caches()
    assert G._engine is None and "_train_engine" not in G.__dict__
    for bad in ("fp32", "BF16", ""):
        with pytest.raises(ValueError):
            G.precision = bad
    G.precision = "bf16"
    assert copy.deepcopy(G).precision == "bf16"
    ada = importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator")
    GA = ada.Generator(512, 512, 2, 16, 3)
    with pytest.raises(ValueError):
        GA.precision = "half"
    GA.synthesis.__dict__["_engine"] = FakeEngine()
    assert "_engine" not in copy.deepcopy(GA).synthesis.__dict__


def test_engine_routes_need_cuda_tensors(sg2):
    G = sg2.Generator(16, 512, 2)
    G.precision = "bf16"
    for p in G.parameters():
        p.requires_grad_(False)
    lat = torch.zeros(1, G.n_latent, 512, requires_grad=True)
    assert not G._use_train_engine(lat, [None] * G.num_layers, False)     # CPU tensors never reach an engine ...
    assert not G._use_engine(lat.detach(), [None] * G.num_layers, False)
    with pytest.raises(RuntimeError, match="CUDA"):                       # ... and the module path refuses them (no CPU fallback)
        G([lat], input_is_latent=True, randomize_noise=False)


def test_engine_holds_its_module_weakly(sg2):
    """module -> engine is the only strong edge: an engine (plan, workspace, CUDA graphs) dies with its module by reference
    counting, never in a cyclic-GC pass that might run inside somebody's stream capture"""
    import gc
    E = importlib.import_module("stylegan-for-facerec_b200.engine")
    G = sg2.Generator(16, 32, 2)
    eng = E.SynthesisEngine.__new__(E.SynthesisEngine)
    eng.plan = None
    eng.G = G
    assert eng.G is G
    assert not any(r is eng or r is getattr(eng, "__dict__", None) for r in gc.get_referrers(G) if isinstance(r, (dict, E.SynthesisEngine)))
    del G
    with pytest.raises(RuntimeError, match="no longer exists"):
        eng.G
    A = sg2.stylegan2_ada.Generator(32, 32, 2, 16, 3)
    aeng = E.AdaSynthesisEngine.__new__(E.AdaSynthesisEngine)
    aeng.plan = None
    aeng.G = E._AdaView(A.synthesis)
    assert aeng.G.syn is A.synthesis
    del A
    with pytest.raises(RuntimeError, match="no longer exists"):
        aeng.G.syn
