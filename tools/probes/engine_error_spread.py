#!/usr/bin/env python
"""How the bf16 engine's error against the exact fp32 path spreads over random latents (per sample: max|d| / max|ref|), at the two
benchmark configurations.  -> profiles/engine_error_spread_r02.json"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

sg2 = importlib.import_module("stylegan-for-facerec_b200")
dev = "cuda:0"
out = {}
for size, n, bs in ((256, 256, 32), (1024, 32, 4)):
    G = bench.make_generator(sg2, size, dev, "bf16")
    g = torch.Generator().manual_seed(99)
    rels = []
    with torch.no_grad():
        for i in range(n // bs):
            z = torch.randn(bs, 512, generator=g).to(dev)
            G.precision = "bf16"
            y = G([z], randomize_noise=False)[0].float()
            G.precision = "exact"
            r = G([z], randomize_noise=False)[0]
            d = (y - r).abs().flatten(1).max(1).values / r.abs().flatten(1).max(1).values
            rels += d.tolist()
            rms = ((y - r).pow(2).flatten(1).mean(1).sqrt() / r.pow(2).flatten(1).mean(1).sqrt()).tolist()
    rels.sort()
    out[str(size)] = {"samples": len(rels), "rel_max_min": round(rels[0], 5), "rel_max_median": round(rels[len(rels) // 2], 5),
                      "rel_max_p90": round(rels[int(0.9 * len(rels))], 5), "rel_max_max": round(rels[-1], 5),
                      "rel_rms_last_batch_median": round(sorted(rms)[len(rms) // 2], 5)}
    del G
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
