#!/usr/bin/env python
"""Golden vectors of the rosinality Discriminator (backbone/stylegan2/model.py:545-673 of the reference), produced by
the UNMODIFIED reference class on CPU with the op shim of make_golden.py.  The reference repo never instantiates it
(SURVEY.md section 8f-4); it is mirrored so that third-party rosinality training code keeps importing, and these vectors
pin that mirror (tests/test_abi_cpu.py runs it on CPU with the oracle's ops patched in, tests/test_model_gpu.py on the
kernels).  Weights are regenerated from name-keyed seeds on both sides (21 M parameters are not committed).

    python tests/golden/make_golden_disc.py        -> tests/golden/disc.npz, disc_state_dict_keys.json
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import sg2_oracle as O  # noqa: E402

# name, size, channel_multiplier, batch
CASES = [("d16", 16, 2, 4), ("d32_b8", 32, 2, 8), ("d64_cm1", 64, 1, 2)]


def seeded_state_dict(module, tag):
    """every floating tensor of module.state_dict() from its own name-keyed stream; biases perturbed away from 0,
    the Blur kernels (buffers) kept"""
    sd = {}
    for k, v in module.state_dict().items():
        if k.endswith(".kernel"):
            sd[k] = v.clone()
        elif k.endswith(".bias"):
            sd[k] = 0.1 * O.named_randn(f"{tag}:{k}", tuple(v.shape), 7)
        else:
            sd[k] = O.named_randn(f"{tag}:{k}", tuple(v.shape), 7)
    return sd


def images(name, batch, size):
    return O.named_randn(name + ":img", (batch, 3, size, size), 8)


def main():
    import make_golden as MG
    M, *_ = MG.import_reference()
    out, keys = {}, {}
    torch.set_grad_enabled(False)
    for name, size, cm, batch in CASES:
        D = M.Discriminator(size, channel_multiplier=cm).eval()
        keys[name] = [(k, list(v.shape)) for k, v in D.state_dict().items()]
        D.load_state_dict(seeded_state_dict(D, name), strict=True)
        y = D(images(name, batch, size))
        out[name + "/out"] = y.numpy()
        print(name, tuple(y.shape), y.flatten()[:4].tolist())
    np.savez_compressed(os.path.join(HERE, "disc.npz"), **out)
    json.dump(keys, open(os.path.join(HERE, "disc_state_dict_keys.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
