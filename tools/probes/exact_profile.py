#!/usr/bin/env python
"""torch.profiler kernel table of one exact-path (fp32) forward -- where the default route spends its time."""
import importlib, os, sys
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
torch.manual_seed(0)
size, B = int(os.environ.get("SIZE", "256")), int(os.environ.get("BATCH", "32"))
G = sg2.Generator(size, 512, 8).to("cuda:0").eval()
G.precision = "exact"
z = torch.randn(B, 512, device="cuda:0")
with torch.no_grad():
    for _ in range(2):
        G([z], randomize_noise=False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        G([z], randomize_noise=False)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))
