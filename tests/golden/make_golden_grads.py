#!/usr/bin/env python
"""Golden GRADIENTS of the whole decoder, produced by autograd through the UNMODIFIED reference modules on CPU
(fp64): d(sum(image * gy)) / d(latents, noise) for the rosinality Generator (backbone/stylegan2/model.py, ops rebound
to the reference's own PyTorch statements as in make_golden.py) and d/d(ws) for the stylegan2_ada Generator.
They pin the oracle's autograd (tests/test_oracle_golden.py), which in turn is what the `-m gpu` gradient tests
compare the CUDA path with.  Runs only in the authoring container (needs /root/reference).

    python tests/golden/make_golden_grads.py        -> tests/golden/grads.npz
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import sg2_ada_oracle as A  # noqa: E402
from oracle import sg2_oracle as O  # noqa: E402

# name, size, n_mlp, batch
ROS_CASES = [("ros16", 16, 2, 2), ("ros32", 32, 2, 1)]
# name, resolution, mapping layers, batch
ADA_CASES = [("ada16", 16, 2, 3), ("ada32", 32, 2, 2)]


def ros_inputs(name, size, batch):
    log_size = int(np.log2(size))
    n_latent, num_layers = 2 * log_size - 2, 2 * (log_size - 2) + 1
    lat = 0.5 * O.named_randn(name + ":lat", (batch, n_latent, 512), 3)
    noise = [O.named_randn(f"{name}:n{i}", (batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 3) for i in range(num_layers)]
    gy = O.named_randn(name + ":gy", (batch, 3, size, size), 3)
    return lat, noise, gy


def ada_inputs(name, res, batch):
    ws = O.named_randn(name + ":ws", (batch, A.num_ws(res), 512), 5)
    gy = O.named_randn(name + ":gy", (batch, 3, res, res), 6)
    return ws, gy


def main():
    import make_golden as MG
    M, *_ = MG.import_reference()
    out = {}
    for name, size, n_mlp, batch in ROS_CASES:
        sd = O.init_state_dict(size, 512, n_mlp, 2, seed=0)
        G = M.Generator(size, 512, n_mlp)
        G.load_state_dict(sd, strict=True)
        G = G.double().eval()
        lat, noise, gy = ros_inputs(name, size, batch)
        ld = lat.double().requires_grad_(True)
        nd = [n.double().requires_grad_(True) for n in noise]
        img, _ = G([ld], input_is_latent=True, noise=nd)
        grads = torch.autograd.grad(img, [ld] + nd, gy.double())
        out[name + "/image"] = img.detach().numpy()
        out[name + "/g_latent"] = grads[0].numpy()
        for i, g in enumerate(grads[1:]):
            out[f"{name}/g_noise{i}"] = g.numpy()
        print(name, "image", tuple(img.shape), "|g_latent|max", grads[0].abs().max().item())
    sys.path.insert(0, "/root/reference/restyle-encoder")
    from models.stylegan2_ada.generator import Generator as AdaGenerator
    for name, res, nl, batch in ADA_CASES:
        sd = A.init_state_dict(res, 512, 512, nl, seed=0)
        with contextlib.redirect_stdout(io.StringIO()):
            G = AdaGenerator(512, 512, nl, res, 3)
        G.load_state_dict(sd, strict=True)
        G = G.double().eval()
        ws, gy = ada_inputs(name, res, batch)
        wd = ws.double().requires_grad_(True)
        img, _ = G([wd], input_is_latent=True, randomize_noise=False)
        (g,) = torch.autograd.grad(img, [wd], gy.double())
        out[name + "/image"] = img.detach().numpy()
        out[name + "/g_ws"] = g.numpy()
        print(name, "image", tuple(img.shape), "|g_ws|max", g.abs().max().item())
    path = os.path.join(HERE, "grads.npz")
    np.savez_compressed(path, **out)
    print("wrote", len(out), "arrays,", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
