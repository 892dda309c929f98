#!/usr/bin/env python
"""Summarise an ncu report's source page: per kernel, the spin counts of every mbarrier wait and the
hottest SASS lines by stall samples.   python tools/ncu_hot.py report.ncu-rep [kernel_index] [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    only = int(sys.argv[2]) if len(sys.argv) > 2 else None
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for ki, k in enumerate(kernels):
        if only is not None and ki != only:
            continue
        hdr, data = k["rows"][0], k["rows"][1:]
        ia, isrc, isamp, iex = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed"))
        tot = sum(int(r[isamp] or 0) for r in data)
        print(f"== kernel {ki}: {k['name'][:60]}  samples={tot} inst={sum(int(r[iex] or 0) for r in data)}")
        for n, r in enumerate(data):
            if "TRYWAIT" in r[isrc] and int(r[iex] or 0) > 0:
                print(f"   wait  {r[ia][-5:]} exec={r[iex]:>10s} samples={r[isamp]:>6s} {r[isrc][30:80]}")
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:top_n]:
            print(f"   {r[isamp]:>7s} {r[iex]:>10s} {r[ia][-5:]} {r[isrc][:100]}")


if __name__ == "__main__":
    main()
