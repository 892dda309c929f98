#!/usr/bin/env python
"""Op microbench sweep (BASELINE.json config 5): upfirdn2d {blur, up2, down2} and fused_leaky_relu /
noise+bias+act across resolutions and channel widths, achieved GB/s vs the measured HBM peak.

Algorithmic bytes: lrelu 2*N*s + C*s; upfirdn2d (N_in + N_out)*s + 64;
noise-add 2*N*s + B*r^2*s.  Timing: CUDA events, 5 warm-up + 20 timed launches; every tensor is
>= 2x L2 where the shape allows (inputs larger than L2), otherwise an L2 flush buffer is written
between launches.   python tools/opbench.py [--quick] > profiles/opbench_rNN.jsonl
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def load_ref_ext():
    """the reference's own CUDA extensions (oracle/_ref, compiled from /root/reference/backbone/stylegan2/op for sm_100a by
    oracle/build_ref.py): timed beside ours on the same tensors where they support the dtype (fp32, fp16)"""
    import importlib.util
    mods = {}
    for name in ("fused", "upfirdn2d"):
        path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location(name, path)
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    return mods


def timeit(fn, flush, iters=20, warm=5):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-ref", action="store_true", help="do not time the reference's own CUDA extensions beside ours")
    ap.add_argument("--res", type=int, nargs="*", default=None, help="only these resolutions")
    ap.add_argument("--dtypes", nargs="*", default=None, help="only these dtypes (float32 bfloat16 float16)")
    args = ap.parse_args()
    import time
    from bench import ClockSampler
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    dev = "cuda:0"
    peak, which = peak_gbs()
    ref = None if args.no_ref else load_ref_ext()
    sampler = ClockSampler(0).start()        # nvidia-smi clocks / throttle reasons, one record per printed line
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)   # 256 MiB > 126 MB L2
    taps = (sg2.make_kernel([1, 3, 3, 1])).to(dev)
    res_list = [4, 16, 64, 256, 1024] if args.quick else [4, 8, 16, 32, 64, 128, 256, 512, 1024]
    ch_list = [3, 64, 512] if args.quick else [3, 32, 64, 128, 256, 512]
    if args.res:
        res_list = list(args.res)
    target_elems = 1 << 28 if not args.quick else 1 << 26     # ~256 Mi elements per tensor
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        if args.dtypes and str(dtype).split(".")[1] not in args.dtypes:
            continue
        if dtype == torch.float16 and ref is None and not args.dtypes:
            continue                          # fp16 only as the 2-byte comparison against the reference kernels
        s = torch.finfo(dtype).bits // 8
        with_ref = ref is not None and dtype in (torch.float32, torch.float16)
        for r in res_list:
            for c in ch_list:
                B = max(1, min(4096, target_elems // (c * r * r)))
                n = B * c * r * r
                if n * s > 6 << 30:
                    continue
                x = torch.randn(B, c, r, r, device=dev, dtype=dtype)
                bias = torch.randn(c, device=dev, dtype=dtype)
                need_flush = flush if n * s < (256 << 20) else None
                rows = []
                empty = x.new_empty(0)
                taps32 = taps.float()

                def ref_up(xx, k, up, down, pad):
                    b_, c_, h_, w_ = xx.shape
                    return ref["upfirdn2d"].upfirdn2d(xx.reshape(-1, h_, w_, 1), k.to(xx.dtype), up, up, down, down, pad[0], pad[1], pad[0], pad[1])

                def run(op, nbytes, ours, theirs):
                    t0 = time.perf_counter()
                    t = timeit(ours, need_flush)
                    tr = timeit(theirs, need_flush, iters=10, warm=3) if (with_ref and theirs is not None) else None
                    torch.cuda.synchronize()
                    rows.append((op, nbytes, t, tr, sampler.summary(t0, time.perf_counter())))

                run("fused_leaky_relu", 2 * n * s + c * s, lambda: sg2.fused_leaky_relu(x, bias),
                    lambda: ref["fused"].fused_bias_act(x, bias, empty, 3, 0, 0.2, 2 ** 0.5))
                if r >= 4 and r <= 512:
                    run("upfirdn2d_up2", (n + 4 * n) * s + 64, lambda: sg2.upfirdn2d(x, taps * 4, up=2, pad=(2, 1)),
                        lambda: ref_up(x, taps32 * 4, 2, 1, (2, 1)))
                if r >= 8:
                    run("upfirdn2d_down2", (n + n // 4) * s + 64, lambda: sg2.upfirdn2d(x, taps, down=2, pad=(1, 1)),
                        lambda: ref_up(x, taps32, 1, 2, (1, 1)))
                xb = torch.randn(B, c, r + 1, r + 1, device=dev, dtype=dtype)
                run("upfirdn2d_blur", (xb.numel() + n) * s + 64, lambda: sg2.upfirdn2d(xb, taps * 4, pad=(1, 1)),
                    lambda: ref_up(xb, taps32 * 4, 1, 1, (1, 1)))
                del xb
                for op, nbytes, t, tr, clk in rows:
                    line = {"op": op, "dtype": str(dtype).split(".")[1], "B": B, "C": c, "res": r,
                            "bytes": nbytes, "ms": round(t * 1e3, 4), "GBps": round(nbytes / t / 1e9, 1),
                            "frac_of_hbm_peak": round(nbytes / t / 1e9 / peak, 3), "peak": which,
                            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "samples": clk["samples"],
                                       "reasons": clk["reasons"]}}
                    if tr is not None:
                        line["ref_ext_ms"] = round(tr * 1e3, 4)
                        line["ref_ext_frac_of_hbm_peak"] = round(nbytes / tr / 1e9 / peak, 3)
                        line["speedup_vs_ref_ext"] = round(tr / t, 2)
                    print(json.dumps(line), flush=True)
                del x
                torch.cuda.empty_cache()
    sampler.stop()


if __name__ == "__main__":
    main()
