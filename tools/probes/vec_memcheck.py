#!/usr/bin/env python
"""A few launches of every new row-fetch path (vector rows of the streaming kernel: fp32 down-2, bf16 / fp16 up-2; the packed
kernel incl. its 2- and 4-lane strips) on tensors that END an allocation, for compute-sanitizer --tool memcheck."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
dev = "cuda:0"
g = torch.Generator().manual_seed(3)
taps = torch.randn(4, 4, generator=g).to(dev)


def tail_tensor(shape, dtype):
    n = 1
    for s in shape:
        n *= s
    big = torch.zeros(n, device=dev, dtype=dtype)          # exactly the tensor: its end is the end of the allocation
    big.copy_(torch.randn(n, generator=g).to(dtype))
    return big.view(shape)


n = 0
for dtype in (torch.float32, torch.bfloat16, torch.float16):
    for shape, up, down, pad in (((2, 3, 64, 64), 1, 2, (1, 1)), ((1, 2, 40, 24), 1, 2, (2, 2)), ((3, 2, 33, 16), 1, 2, (0, 1)), ((1, 2, 70, 300), 1, 2, (3, 1)),
                                 ((2, 3, 40, 72), 2, 1, (2, 1)), ((1, 2, 33, 256), 2, 1, (2, 2)), ((3, 5, 20, 8), 2, 1, (2, 1)), ((1, 3, 17, 16), 2, 1, (2, 0)),
                                 ((2, 3, 65, 65), 1, 1, (1, 1)), ((1, 2, 257, 257), 1, 1, (1, 1)), ((9473, 1, 17, 17), 1, 1, (1, 1)), ((9500, 1, 33, 33), 1, 1, (1, 1)),
                                 ((9473, 1, 16, 32), 1, 2, (1, 1)), ((19000, 1, 16, 16), 1, 2, (1, 1)), ((19000, 1, 17, 17), 1, 1, (2, 2))):
        x = tail_tensor(shape, dtype)
        y = sg2.upfirdn2d(x, taps, up, down, pad)
        assert torch.isfinite(y.float()).all()
        n += 1
torch.cuda.synchronize()
print("launched", n)
