#!/bin/bash
# bulk-copy loader of the packed kernel: correctness (tight timeouts: a hang must not eat the budget), then timing
mkdir -p gpurun_out/pk14
timeout 90 python tools/probes/pk_check.py --quick > gpurun_out/pk14/quick.txt 2>&1; echo "quick rc=$?"; tail -2 gpurun_out/pk14/quick.txt | cut -c1-200
timeout 120 python tools/probes/pk_check.py --perf > gpurun_out/pk14/check.txt 2> gpurun_out/pk14/check.err; echo "check rc=$?"
grep -c MISMATCH gpurun_out/pk14/check.txt; grep "^checked" gpurun_out/pk14/check.txt | cut -c1-200
grep bfloat16 gpurun_out/pk14/check.txt | grep '"op"' | cut -c1-120
