#!/usr/bin/env python
"""Medians of an op-sweep file (tools/opbench.py output) per op and dtype over the planes >= MIN_RES (default 64): fraction of the
measured HBM peak (median, min, max), median speed-up against the reference's own kernels where they were timed, and how many
lines carry a slowdown reason.     python tools/opbench_summary.py profiles/opbench_r02_v3.jsonl [min_res]"""
import collections
import json
import statistics
import sys


def main():
    path = sys.argv[1]
    min_res = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rows = [json.loads(l) for l in open(path) if l.strip().startswith("{")]
    frac, speed = collections.defaultdict(list), collections.defaultdict(list)
    for r in rows:
        if r.get("res", 0) >= min_res:
            frac[(r["op"], r["dtype"])].append(r["frac_of_hbm_peak"])
            if "speedup_vs_ref_ext" in r:
                speed[(r["op"], r["dtype"])].append(r["speedup_vs_ref_ext"])
    for k in sorted(frac):
        v = frac[k]
        sp = f"  {statistics.median(speed[k]):.2f}x the reference kernel" if speed[k] else ""
        print(f"{k[0]:18s} {k[1]:9s} median {statistics.median(v):.3f}  min {min(v):.3f}  max {max(v):.3f}  ({len(v)} shapes){sp}")
    slow = sum(1 for r in rows if any("slowdown" in x for x in r.get("clocks", {}).get("reasons", [])))
    print(f"{len(rows)} lines, {slow} with a slowdown reason")


if __name__ == "__main__":
    main()
