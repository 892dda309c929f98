#!/usr/bin/env python
"""Randomised differential test of upfirdn2d over every dispatch path: random plane counts / sizes / taps / pads / dtypes for the
three model geometries against the fp64 oracle (small tensors), and many-plane tensors (the thresholds that switch the packed
kernel to 2 / 4 lanes per strip and to planes-8-apart grouping) against the generic tiled kernels (SG2_UPFIRDN_* switches off)."""
import importlib
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sg2 = importlib.import_module("stylegan-for-facerec_b200")
import sg2_oracle as oracle  # noqa: E402

DEV = "cuda:0"
TOL = {torch.float32: 2e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}
rng = random.Random(int(os.environ.get("SEED", "7")))
g = torch.Generator().manual_seed(rng.randrange(1 << 30))
bad = n = 0
for it in range(int(os.environ.get("CASES", "400"))):
    dtype = rng.choice([torch.float32, torch.bfloat16, torch.float16])
    up, down = rng.choice([(1, 1), (2, 1), (1, 2)])
    k = rng.choice([4, 4, 4, 3, 2])
    h, w = rng.choice([rng.randint(5, 40), rng.randint(40, 140), rng.choice([16, 32, 64, 128, 33, 65, 129, 257])]), \
        rng.choice([rng.randint(5, 40), rng.randint(40, 300), rng.choice([16, 32, 64, 128, 256, 33, 65, 129, 257, 264])])
    pad = (rng.randint(0, 3), rng.randint(0, 3)) if rng.random() < 0.7 else {(1, 1): (1, 1), (2, 1): (2, 1), (1, 2): (1, 1)}[(up, down)]
    planes = rng.choice([1, 2, 3, 5, 8, 13])
    if h * up + pad[0] + pad[1] < k or w * up + pad[0] + pad[1] < k:
        continue
    sep = rng.random() < 0.4
    taps = torch.outer(torch.randn(k, generator=g), torch.randn(k, generator=g)) if sep else torch.randn(k, k, generator=g)
    x = torch.randn(1, planes, h, w, generator=g).to(dtype)
    y = sg2.upfirdn2d(x.to(DEV), taps.to(DEV), up, down, pad)
    ref = oracle.upfirdn2d(x.double(), taps.double(), up, down, pad)
    n += 1
    err = float((y.double().cpu() - ref).abs().max())
    lim = TOL[dtype] * max(float(ref.abs().max()), 1e-6) * (k if dtype == torch.float32 else 1)
    if y.shape != ref.shape or not (err <= lim):
        bad += 1
        print(f"MISMATCH {dtype} planes {planes} {h}x{w} k {k} up {up} down {down} pad {pad} sep {sep}: err {err:.3g} > {lim:.3g}", flush=True)
# many planes: the dispatch thresholds of the packed kernel (2 / 4 lanes per strip, planes 8 apart) against the tiled kernels
for dtype in (torch.bfloat16, torch.float16):
    for (planes, h, w, down, pad) in ((19001, 17, 17, 1, (1, 1)), (19003, 16, 16, 2, (1, 1)), (9477, 33, 33, 1, (1, 1)), (9479, 32, 32, 2, (1, 1)),
                                      (4801, 65, 65, 1, (1, 1)), (4803, 64, 64, 2, (1, 1)), (2501, 129, 129, 1, (1, 1)), (9481, 20, 24, 1, (2, 2)),
                                      (19005, 12, 18, 1, (0, 2)), (9483, 40, 34, 2, (2, 2))):
        taps = torch.randn(4, 4, generator=g).to(DEV)
        x = torch.randn(1, planes, h, w, generator=g).to(dtype).to(DEV)
        y = sg2.upfirdn2d(x, taps, 1, down, pad)
        for key in ("SG2_UPFIRDN_PK", "SG2_UPFIRDN_PLANES"):
            os.environ[key] = "0"
        y0 = sg2.upfirdn2d(x, taps, 1, down, pad)
        for key in ("SG2_UPFIRDN_PK", "SG2_UPFIRDN_PLANES"):
            del os.environ[key]
        n += 1
        d = (y.float() - y0.float()).abs()
        lim = (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10) * torch.maximum(y.float().abs(), y0.float().abs()) + 1e-5 * float(y0.float().abs().max())
        if not bool((d <= lim).all()):
            bad += 1
            print(f"MANY-PLANES MISMATCH {dtype} planes {planes} {h}x{w} down {down} pad {pad}: max {float(d.max()):.4g}", flush=True)
print(f"fuzz: {n} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
