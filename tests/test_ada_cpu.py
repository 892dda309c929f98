"""CPU: the ADA-decoder oracle is pinned to golden vectors produced by the unmodified reference module
(tests/golden/make_golden_ada.py), and the product's module tree has the reference's state_dict keys."""
import importlib
import json
import os

import numpy as np
import torch

from conftest import GOLDEN


def _golden():
    return np.load(os.path.join(GOLDEN, "ada.npz"))


def test_ada_oracle_matches_reference():
    import make_golden_ada as M
    from oracle import sg2_ada_oracle as A
    from oracle.sg2_oracle import named_randn
    g = _golden()
    for name, res, nl, b, mode, (psi, cutoff) in M.CASES:
        if res > 64:
            continue                      # keep the CPU suite short; the 128^2 case runs in the GPU test
        sd = A.init_state_dict(res, 512, 512, nl, seed=0)
        if mode == "z":
            z = named_randn("ada:z:" + name, (b, 512), 1)
            ws = A.mapping_network(sd, z, A.num_ws(res), nl, psi, cutoff)
            assert torch.equal(ws, torch.from_numpy(g[name + "/ws"])), name
            img = A.synthesis_network(sd, res, ws)
        else:
            ws = named_randn("ada:w:" + name, (b, A.num_ws(res), 512), 1)
            img = A.generator_forward(sd, res, ws, nl, input_is_latent=True)
        assert torch.equal(img, torch.from_numpy(g[name + "/image"])), name     # same library calls -> bit-exact
    for name, shape in (("up_5x7", (2, 3, 5, 7)), ("up_1x1", (1, 2, 1, 1)), ("up_16", (1, 4, 16, 16))):
        x = named_randn("ada:up:" + name, shape, 2)
        assert torch.equal(A.smooth_upsample(x, A.smooth_kernel()), torch.from_numpy(g["smooth_upsample/" + name])), name


def test_ada_module_tree_has_the_reference_state_dict_keys(sg2):
    ada = importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator")
    ref_keys = json.load(open(os.path.join(GOLDEN, "ada_state_dict_keys.json")))
    for res, nl in ((32, 2), (64, 8), (128, 2)):
        G = ada.Generator(512, 512, nl, res, 3)
        want = ref_keys[str(res)]
        mine = {k: list(v.shape) for k, v in G.state_dict().items()}
        if res == 64:
            want = {k: v for k, v in want.items() if not k.startswith("mapping.layers")}   # golden 64 has 8 and 2 layers
            mine = {k: v for k, v in mine.items() if not k.startswith("mapping.layers")}
        assert mine == want, res
        assert G.num_ws == 2 * (int(np.log2(res)) - 1 + 1)
    from oracle import sg2_ada_oracle as A
    G = ada.Generator(512, 512, 2, 32, 3)
    G.load_state_dict(A.init_state_dict(32, 512, 512, 2), strict=True)


def test_ada_cpu_tensor_raises(sg2):
    ada = importlib.import_module("stylegan-for-facerec_b200.stylegan2_ada.generator")
    G = ada.Generator(512, 512, 2, 32, 3).eval()
    try:
        with torch.no_grad():
            G([torch.randn(1, 512)], randomize_noise=False)
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("a CPU tensor must raise like the reference's CHECK_CUDA")
