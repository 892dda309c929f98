#!/usr/bin/env python
"""Compare per-kernel CUDA-event tables written by bench.py --profile-out.
   python tools/kcmp.py [--kind gemm|upfir|...] [--min-ms 0.02] a.json b.json ..."""
import argparse
import json


def load(path):
    d = json.load(open(path))
    return d["kernels"] if isinstance(d, dict) else d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--kind", default=None)
    ap.add_argument("--min-ms", type=float, default=0.02)
    a = ap.parse_args()
    D = [load(f) for f in a.files]
    print(" " * 40 + "".join(f"{f.split('/')[-1][:11]:>12s}" for f in a.files))
    for i, k in enumerate(D[0]):
        if (a.kind and k["kind"] != a.kind) or k["s"] * 1e3 < a.min_ms:
            continue
        extra = ""
        s_last = D[-1][i]["s"]
        if k["kind"] == "gemm" and s_last > 0:
            extra = f"  {k['flops'] / s_last / 1e12:6.0f} TF/s"
        if k["kind"] == "upfir" and s_last > 0:
            extra = f"  {k['bytes'] / s_last / 1e9:6.0f} GB/s"
        print(f"{k['kind'][:5]:6s}{k['what']:34s}" + "".join(f"{d[i]['s'] * 1e3:12.3f}" for d in D) + extra)
    for kind in ("gemm", "upfir", "rgb_combine"):
        print(f"{kind + ' total':40s}" + "".join(f"{sum(k['s'] for k in d if k['kind'] == kind) * 1e3:12.3f}" for d in D))
    print(f"{'all total':40s}" + "".join(f"{sum(k['s'] for k in d) * 1e3:12.3f}" for d in D))


if __name__ == "__main__":
    main()
