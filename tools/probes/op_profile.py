#!/usr/bin/env python
"""A handful of op launches for `ncu --set full` (upfirdn2d blur / down-2 / up-2 at fp32 and bf16, 256^2 planes).
   ncu --set full --clock-control none --import-source on -k regex:upfirdn2d -o rep python tools/probes/op_profile.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
dev = "cuda:0"
taps = sg2.make_kernel([1, 3, 3, 1]).to(dev)
res = int(os.environ.get("OP_RES", "256"))
for dtype in (torch.float32, torch.bfloat16):
    x = torch.randn(64, 64, res, res, device=dev, dtype=dtype)
    xb = torch.randn(64, 64, res + 1, res + 1, device=dev, dtype=dtype)
    for _ in range(2):
        sg2.upfirdn2d(xb, taps * 4, pad=(1, 1))
        sg2.upfirdn2d(x, taps, down=2, pad=(1, 1))
        sg2.upfirdn2d(x, taps * 4, up=2, pad=(2, 1))
    torch.cuda.synchronize()
print("done")
