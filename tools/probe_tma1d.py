#!/usr/bin/env python
"""probe: which 1-D tensor maps cuTensorMapEncodeTiled accepts (through the library's own entry point)"""
import ctypes as C, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sg2 = importlib.import_module("stylegan-for-facerec_b200")
lib = sg2._lib.load()
f = lib.sg2_debug_probe_tma1d
f.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int]
f.restype = C.c_int
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
for es in (4, 2):
    for box in (20, 24, 36, 68, 72, 132, 136):
        if (box * es) % 16:
            continue
        row = []
        for lg in (8, 12, 15, 18, 20, 21, 22, 23, 24, 25, 26, 28):
            row.append("%d:%d" % (lg, f(x.data_ptr(), 1 << lg, box, es)))
        print("es", es, "box", box, " ".join(row), flush=True)
for n in (32768, 8405000, 8388608, 1060900, 3145728, 203574):
    print("n", n, [f(x.data_ptr(), n, 68, 4), f(x.data_ptr(), n, 72, 2), f(x.data_ptr() + 16, n, 68, 4)])
