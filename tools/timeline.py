#!/usr/bin/env python
"""Kernel timeline of a few bench steps (torch.profiler / CUPTI): every GPU activity with its start,
duration and the idle gap before it, so time spent OUTSIDE the engine's kernels (mapping, staging
copies, graph launch gaps) shows up.  Diagnostic only -- numbers under a profiler are not bench values.

    python tools/timeline.py --size 256 --batch 64 --steps 3 [--e2e] > gpurun_out/timeline.txt
"""
import argparse
import importlib
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--e2e", action="store_true", help="host latents in, images back to pinned host memory")
    args = ap.parse_args()
    import bench
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    dev = torch.device("cuda:0")
    G = bench.make_generator(sg2, args.size, dev, "bf16")
    z_host = torch.randn(args.steps + 4, args.batch, 512).pin_memory()
    z_dev = z_host.to(dev)
    out_host = torch.empty(args.batch, 3, args.size, args.size).pin_memory()

    def step(i):
        with torch.no_grad():
            if args.e2e:
                im = G([z_host[i].to(dev, non_blocking=True)], randomize_noise=False)[0]
                out_host.copy_(im, non_blocking=True)
            else:
                im = G([z_dev[i]], randomize_noise=False)[0]
        return im

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            step(4 + i)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    prev_end = t0
    busy = 0.0
    print(f"# {len(evs)} GPU activities over {args.steps} steps; times in us")
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        busy += d
        print(f"{s:10.1f} {d:9.1f} gap {gap:7.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)
    span = prev_end - t0
    print(f"# span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us ({100 * (span - busy) / span:.1f} %), "
          f"{span / args.steps:.1f} us per step")


if __name__ == "__main__":
    main()
