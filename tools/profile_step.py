#!/usr/bin/env python
"""One synthesis forward inside a cudaProfilerStart/Stop range, for ncu.

    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:'modconv_gemm|upfir_tc' -o gpurun_out/step python tools/profile_step.py --size 256 --batch 64

The forward is warmed up first (plan, descriptors, lazy attributes) and runs without a CUDA graph so
every kernel is an ordinary launch.  Numbers printed under a profiler are never bench values.
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SG2_B200_GRAPH", "0")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    G = sg2.Generator(args.size, 512, 8).to(dev).eval()
    G.precision = "bf16"
    z = torch.randn(args.batch, 512, device=dev)
    with torch.no_grad():
        for _ in range(2):
            G([z])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        G([z])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("profiled one forward", args.size, args.batch)


if __name__ == "__main__":
    main()
