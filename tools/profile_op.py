#!/usr/bin/env python
"""One upfirdn2d call per geometry inside a cudaProfilerStart/Stop range, for ncu (numbers under a profiler are not bench values).
    ncu --profile-from-start off --set full --import-source on -k regex:upfirdn2d_stream -o gpurun_out/op python tools/profile_op.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
dev = "cuda:0"
taps = sg2.make_kernel([1, 3, 3, 1]).to(dev)
cases = []
for dt in (torch.float32, torch.bfloat16):
    cases.append((torch.randn(16, 64, 257, 257, device=dev, dtype=dt), dict(pad=(1, 1))))
    cases.append((torch.randn(16, 64, 256, 256, device=dev, dtype=dt), dict(down=2, pad=(1, 1))))
    cases.append((torch.randn(16, 64, 128, 128, device=dev, dtype=dt), dict(up=2, pad=(2, 1))))
for x, kw in cases:
    sg2.upfirdn2d(x, taps, **kw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for x, kw in cases:
    sg2.upfirdn2d(x, taps, **kw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", len(cases))
