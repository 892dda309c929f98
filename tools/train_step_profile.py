"""torch.profiler kernel table of one decoder forward + backward through the differentiable tensor-core route
(which passes are left after the fused layers).   python tools/train_step_profile.py [--batch 32] [--ada]"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sg2_b200 as sg2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ada", action="store_true")
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
    G.precision = "bf16"
    lat = torch.randn(a.batch, n_latent, 512, device=dev)
    gy = torch.randn(a.batch, 3, a.size, a.size, device=dev)

    def step():
        ld = lat.clone().requires_grad_(True)
        img, _ = G([ld], input_is_latent=True, randomize_noise=False)
        img.backward(gy)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))


if __name__ == "__main__":
    main()
