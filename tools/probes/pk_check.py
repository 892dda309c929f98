#!/usr/bin/env python
"""upfirdn2d_pk.cu (packed-math kernel for 2-byte storage) against the row-streaming kernel it replaces: same taps in the
same order with IEEE fused multiply-adds, so the outputs must be IDENTICAL bit for bit.  Shapes cover every alignment q of
the row segment (odd pitches), every lanes-per-group instantiation, several strips / bands / super-bands, tensors placed
at the very end of an allocation behind NaNs (over-reads would show), and both dtypes.  Then a same-box timing A/B.

    python tools/probes/pk_check.py [--perf] > gpurun_out/pk_check.txt
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DEV = "cuda:0"


def run(sg2, x, taps, down, pad, pk):
    os.environ["SG2_UPFIRDN_PK"] = "1" if pk else "0"
    return sg2.upfirdn2d(x, taps, 1, down, pad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--perf-only", action="store_true", help="timing of the packed kernel only (knock-out variant libraries)")
    ap.add_argument("--quick", action="store_true", help="a few shapes only (for a compute-sanitizer run)")
    args = ap.parse_args()
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    g = torch.Generator().manual_seed(11)
    bad = 0
    n = nsep = ndiff = ntot = 0
    shapes = []
    for w in (33, 40, 47, 64, 65, 66, 67, 68, 69, 70, 71, 72, 100, 128, 129, 131, 200, 255, 256, 257, 258, 259, 260, 261, 263, 300, 513, 520, 1025):
        for h in (9, 17, 40, 65, 130):
            if w * h > 200000:
                continue
            shapes.append((3, h, w))
    if args.quick:
        shapes = [(2, 17, 65), (1, 40, 131), (2, 65, 259)]
    shapes += [(19000, 17, 17), (18950, 9, 20), (9500, 33, 33), (9473, 17, 40), (2, 257, 257), (1, 513, 513), (1, 1025, 1025), (5, 129, 129), (4, 300, 131), (1, 1030, 70), (2, 64, 520)]
    for dtype in (() if args.perf_only else (torch.bfloat16, torch.float16)):
        for (pl, h, w) in shapes:
            for down, pad, k in ((1, (1, 1), 4), (1, (2, 2), 4), (1, (0, 2), 3), (1, (3, 0), 4), (1, (-1, 1), 4), (2, (1, 1), 4), (2, (2, 2), 4), (2, (0, 1), 4),
                                 (2, (1, 0), 2)):
                oh = (h + pad[0] + pad[1] - k) // down + 1
                ow = (w + pad[0] + pad[1] - k) // down + 1
                if oh < 1 or ow < 1:
                    continue
                taps = torch.randn(k, k, generator=g).to(DEV)
                numel = pl * h * w
                # the input is the very end of an allocation that is NaN everywhere else, at a 16-byte aligned offset
                big = torch.full((numel + 4096 + 8,), float("nan"), device=DEV, dtype=dtype)
                off = big.numel() - numel
                off -= off % 8                      # 16-byte aligned start; at most 7 elements of NaN tail behind the tensor
                x = big[off:off + numel].view(1, pl, h, w)
                x.copy_(torch.randn(1, pl, h, w, generator=g).to(dtype))
                y1 = run(sg2, x, taps, down, pad, True)
                y0 = run(sg2, x, taps, down, pad, False)
                n += 1
                if ow <= 128 * (8 // (4 * down)) // 2 + 0 and w <= 140:
                    # narrow planes: the groups of a warp as planes 8 apart (chosen by itself only for thousands of planes)
                    os.environ["SG2_UPFIRDN_PK_BYPLANES"] = "1"
                    for npl in (1, 13, 37):
                        xs = torch.randn(1, npl, h, w, generator=g).to(dtype).to(DEV)
                        yb = run(sg2, xs, taps, down, pad, True)
                        yr = run(sg2, xs, taps, down, pad, False)
                        n += 1
                        if not torch.equal(yb.view(torch.int16), yr.view(torch.int16)):
                            bad += 1
                            print(f"BY-PLANES MISMATCH {dtype} planes {npl} {h}x{w} down {down} pad {pad} k {k}", flush=True)
                    del os.environ["SG2_UPFIRDN_PK_BYPLANES"]
                if down == 1:
                    # outer-product taps: the packed kernel takes its separable blur (fp32 rounding order differs): the
                    # outputs may differ from the 16-tap order by one unit in the last place of the storage type
                    ts = torch.outer(torch.randn(k, generator=g), torch.randn(k, generator=g)).to(DEV)
                    z1 = run(sg2, x, ts, down, pad, True).float()
                    z0 = run(sg2, x, ts, down, pad, False).float()
                    eps = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
                    lim = eps * torch.maximum(z0.abs(), z1.abs()) + 4e-6 * float(z0.abs().max())   # floor: outputs that cancel to ~0
                    nsep += 1
                    if not bool(((z1 - z0).abs() <= lim).all()) or not torch.isfinite(z1).all():
                        bad += 1
                        d = (z1 - z0).abs() - lim
                        i = int(d.argmax())
                        print(f"SEPARABLE MISMATCH {dtype} planes {pl} {h}x{w} pad {pad} k {k}: max {float((z1 - z0).abs().max()):.4g}; worst: "
                              f"16-tap {float(z0.flatten()[i])!r} separable {float(z1.flatten()[i])!r} at {i}, taps {ts.flatten().tolist()}", flush=True)
                    ndiff += int((z1 != z0).sum())
                    ntot += z1.numel()
                same = torch.equal(y1.view(torch.int16), y0.view(torch.int16))
                if not same or not torch.isfinite(y1.float()).all():
                    bad += 1
                    d = (y1.float() - y0.float()).abs()
                    idx = torch.nonzero(d > 0)
                    print(f"MISMATCH {dtype} planes {pl} {h}x{w} down {down} pad {pad} k {k}: max {float(d.max()):.4g}, {idx.shape[0]} elements, first {idx[:4].tolist()}",
                          flush=True)
    print(f"checked {n} cases bit for bit + {nsep} with outer-product taps ({ndiff} of {ntot} outputs one ulp away), {bad} mismatches", flush=True)

    if args.perf or args.perf_only:
        peak = 6550.0
        pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pj):
            peak = json.load(open(pj))["hbm_gbs"]
        taps = (sg2.make_kernel([1, 3, 3, 1]) * 4).to(DEV)
        flush = torch.zeros(256 << 20, dtype=torch.uint8, device=DEV)
        for dtype in ((torch.bfloat16,) if args.perf_only else (torch.bfloat16, torch.float16)):
            for r, c in (((256, 128),) if args.perf_only else ((16, 512), (32, 512), (64, 512), (128, 256), (256, 128), (512, 64), (1024, 32))):
                B = max(1, (1 << 28) // (c * r * r))
                for op, down, pad, shape in (("blur", 1, (1, 1), (B, c, r + 1, r + 1)), ("down2", 2, (1, 1), (B, c, r, r))):
                    x = torch.randn(shape, device=DEV, dtype=dtype)
                    oh = (shape[2] + 2 - 4) // down + 1
                    nbytes = (x.numel() + B * c * oh * oh) * 2
                    res = {}
                    for pk in ((True,) if args.perf_only else (False, True)):
                        for _ in range(3):
                            run(sg2, x, taps, down, pad, pk)
                        ts = []
                        for _ in range(10):
                            flush.add_(1)
                            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            a.record()
                            run(sg2, x, taps, down, pad, pk)
                            b.record()
                            torch.cuda.synchronize()
                            ts.append(a.elapsed_time(b))
                        res[pk] = sorted(ts)[len(ts) // 2]
                    line = {"op": op, "dtype": str(dtype).split(".")[1], "res": r, "C": c, "B": B, "pk_ms": round(res[True], 4),
                            "pk_frac": round(nbytes / res[True] / 1e6 / peak, 3)}
                    if False in res:
                        line.update({"stream_ms": round(res[False], 4), "stream_frac": round(nbytes / res[False] / 1e6 / peak, 3)})
                    print(json.dumps(line), flush=True)
                    del x
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
