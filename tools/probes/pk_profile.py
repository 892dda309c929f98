#!/usr/bin/env python
"""Two launches each of the packed-math upfirdn2d kernel (bf16 blur 257^2 -> 256^2 and down-2 256^2 -> 128^2) for ncu.
   ncu --set full --clock-control none --import-source on -k regex:upfirdn2d_pk -o rep python tools/probes/pk_profile.py"""
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
x = torch.randn(64, 64, res, res, device=dev, dtype=torch.bfloat16)
xb = torch.randn(64, 64, res + 1, res + 1, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    sg2.upfirdn2d(xb, taps * 4, pad=(1, 1))
    sg2.upfirdn2d(x, taps, down=2, pad=(1, 1))
torch.cuda.synchronize()
print("done")
