"""Decoder forward + backward (gradient w.r.t. the latents, frozen decoder: the ReStyle direction)
through the differentiable path, fp32 route vs the tensor-core route (precision='bf16').

    python tools/train_step_bench.py [--size 256] [--batch 8] [--iters 10]

Prints one JSON line per mode.  CUDA-event timing, 3 warm-up iterations."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sg2_b200 as sg2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ada", action="store_true", help="the stylegan2_ada decoder instead of the rosinality one")
    a = ap.parse_args()
    dev = "cuda:0"
    torch.manual_seed(0)
    if a.ada:
        G = sg2.stylegan2_ada.Generator(512, 512, 8, a.size, 3).to(dev).eval()
        n_latent = G.num_ws
    else:
        G = sg2.Generator(a.size, 512, 8).to(dev).eval()
        n_latent = G.n_latent
    for p in G.parameters():
        p.requires_grad_(False)
    lat = torch.randn(a.batch, n_latent, 512, device=dev)
    gy = torch.randn(a.batch, 3, a.size, a.size, device=dev)
    # "bf16": the training engine (one autograd node, csrc/synth_train.cu) where it applies, else the layer-by-layer
    # tensor-core route; "bf16-layerwise" forces the latter (SG2_B200_TRAIN_ENGINE=0)
    for mode in ("exact", "bf16-layerwise", "bf16"):
        G.precision = mode.split("-")[0]
        os.environ["SG2_B200_TRAIN_ENGINE"] = "0" if mode == "bf16-layerwise" else "1"
        times = []
        for it in range(3 + a.iters):
            ld = lat.clone().requires_grad_(True)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize()
            e0.record()
            img, _ = G([ld], input_is_latent=True, randomize_noise=False)
            e1.record()
            img.backward(gy)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        with torch.no_grad():                                   # the no-autograd call of the same module, for scale
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                G([lat], input_is_latent=True, randomize_noise=False)
                e1.record()
                torch.cuda.synchronize()
            infer = e0.elapsed_time(e1)
        f = sorted(t[0] for t in times)[len(times) // 2]
        b = sorted(t[1] for t in times)[len(times) // 2]
        print(json.dumps({"mode": mode, "size": a.size, "batch": a.batch, "fwd_ms": round(f, 2), "bwd_ms": round(b, 2),
                          "img_per_s": round(a.batch / (f + b) * 1e3, 1), "no_grad_fwd_ms": round(infer, 2), "decoder": "ada" if a.ada else "rosinality",
                          "grad_norm": float(ld.grad.norm())}), flush=True)


if __name__ == "__main__":
    main()
