#!/usr/bin/env python
"""Golden vectors of the StyleGAN2-ADA decoder variant, produced by the UNMODIFIED reference module
(/root/reference/restyle-encoder/models/stylegan2_ada/generator.py, pure PyTorch -> runs on CPU).

Runs only in the authoring container.  Weights come from oracle.sg2_ada_oracle.init_state_dict (name-keyed
streams) and are loaded with strict=True, which also pins the state_dict key set / shapes.

    python tests/golden/make_golden_ada.py        -> tests/golden/ada.npz, ada_state_dict_keys.json
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import sg2_ada_oracle as A  # noqa: E402
from oracle.sg2_oracle import named_randn  # noqa: E402

CASES = [
    # name, img_resolution, mapping layers, batch, mode, truncation (psi, cutoff)
    ("z_32", 32, 2, 2, "z", (1, None)),
    ("z_64_trunc", 64, 8, 2, "z", (0.7, 4)),
    ("z_64_trunc_all", 64, 2, 1, "z", (0.5, None)),
    ("wplus_64", 64, 2, 3, "w+", (1, None)),
    ("wplus_128", 128, 2, 1, "w+", (1, None)),
]


def main():
    sys.path.insert(0, "/root/reference/restyle-encoder")
    from models.stylegan2_ada.generator import Generator
    out, keys = {}, {}
    torch.set_num_threads(os.cpu_count())
    for name, res, nl, b, mode, (psi, cutoff) in CASES:
        sd = A.init_state_dict(res, 512, 512, nl, seed=0)
        with contextlib.redirect_stdout(io.StringIO()):      # the reference constructor prints a banner
            G = Generator(512, 512, nl, res, 3)
        G.load_state_dict(sd, strict=True)
        G.eval()
        keys[str(res)] = {k: list(v.shape) for k, v in G.state_dict().items()}
        with torch.no_grad():
            if mode == "z":
                z = named_randn("ada:z:" + name, (b, 512), 1)
                ws = G.mapping(z, truncation_psi=psi, truncation_cutoff=cutoff)
                img = G.synthesis(ws, "const")[0]
                out[name + "/ws"] = ws.numpy()
                if psi == 1:
                    img2, none = G([z], randomize_noise=False)                     # the call psp.py makes
                    assert none is None and torch.equal(img, img2)
            else:
                ws = named_randn("ada:w:" + name, (b, A.num_ws(res), 512), 1)
                img, lat = G([ws], input_is_latent=True, randomize_noise=False, return_latents=True)
                assert lat is ws
            ref = A.generator_forward(sd, res, z if mode == "z" else ws, nl, mode != "z", psi, cutoff)
            print(f"{name}: image {tuple(img.shape)} |max| {img.abs().max():.3f}; oracle max|d| {(ref - img).abs().max():.2e}")
            out[name + "/image"] = img.numpy()
    # op-level vectors: SmoothUpsample on odd shapes
    from models.stylegan2_ada.utils import SmoothUpsample
    up = SmoothUpsample()
    for name, shape in (("up_5x7", (2, 3, 5, 7)), ("up_1x1", (1, 2, 1, 1)), ("up_16", (1, 4, 16, 16))):
        x = named_randn("ada:up:" + name, shape, 2)
        with torch.no_grad():
            out["smooth_upsample/" + name] = up(x).numpy()
    np.savez_compressed(os.path.join(HERE, "ada.npz"), **out)
    json.dump(keys, open(os.path.join(HERE, "ada_state_dict_keys.json"), "w"), indent=0, sort_keys=True)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
