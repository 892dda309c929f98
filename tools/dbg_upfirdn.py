#!/usr/bin/env python
"""debug helper: one upfirdn2d call per geometry through the streaming kernel, checked against the oracle"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
from oracle import sg2_oracle as O
torch.manual_seed(0)
cases = [((1, 2, 40, 40), 1, 1, (1, 1)), ((1, 2, 257, 257), 1, 1, (1, 1)), ((1, 2, 64, 64), 2, 1, (2, 1)), ((1, 2, 64, 64), 1, 2, (1, 1)),
         ((1, 1, 300, 300), 1, 1, (2, 2)), ((2, 2, 37, 45), 2, 1, (1, 2))]
if len(sys.argv) > 1:
    cases = cases[int(sys.argv[1]):int(sys.argv[1]) + 1]
for shape, up, down, pad in cases:
    for dt in ((torch.bfloat16, torch.float32) if os.environ.get('DBG_BF16_FIRST') else (torch.float32, torch.bfloat16)):
        x = torch.randn(shape).to(dt)
        k = torch.randn(4, 4)
        y = sg2.upfirdn2d(x.cuda(), k.cuda(), up, down, pad)
        torch.cuda.synchronize()
        ref = O.upfirdn2d(x.double(), k.double(), up, down, pad)
        print(shape, up, down, pad, dt, "max err", float((y.double().cpu() - ref).abs().max()), flush=True)
