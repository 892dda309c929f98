#!/usr/bin/env python
"""Per-launch summary of an ncu report (raw page): duration, DRAM bytes, pipe / memory utilisation.
    python tools/ncu_summary.py gpurun_out/step.ncu-rep > profiles/ncu_full_rNN_summary.json
bench.py reads the committed summary to fill roofline.traffic (DRAM bytes of the dominant kernel per step)."""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[ci["Kernel Name"]].split("(")[0][:60]}
        for k in KEEP:
            if k not in ci or r[ci[k]] == "":
                continue
            v, u = float(r[ci[k]].replace(",", "")), units[ci[k]]
            if u in SCALE and ("byte" in u or u in ("ns", "us", "ms", "s")):
                v *= SCALE[u]
                u = "byte" if "byte" in u else "s"
            d[k] = v if u in ("", "%") else {"value": v, "unit": u}
        res.append(d)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
