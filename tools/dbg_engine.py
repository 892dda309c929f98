#!/usr/bin/env python
"""debug helper: one bf16-engine forward at a given size / batch without CUDA graph, checked against the exact path"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SG2_B200_GRAPH", "0")
sg2 = importlib.import_module("stylegan-for-facerec_b200")
size, batch = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
G = sg2.Generator(size, 512, 2).cuda().eval()
z = torch.randn(batch, 512, device="cuda")
with torch.no_grad():
    G.precision = "bf16"
    a = G([z], randomize_noise=False)[0]
    torch.cuda.synchronize()
    G.precision = "exact"
    b = G([z], randomize_noise=False)[0]
    torch.cuda.synchronize()
print("size", size, "batch", batch, "max rel err", float((a - b).abs().max() / b.abs().max()))
