#!/bin/bash
mkdir -p gpurun_out/pk10
for c in 1 2; do
  echo "== SG2_UPFIRDN_PK_CPL=$c" | tee -a gpurun_out/pk10/cpl.txt
  SG2_UPFIRDN_PK_CPL=$c timeout 200 python tools/probes/pk_check.py --perf-only 2>&1 | grep down2 | tee -a gpurun_out/pk10/cpl.txt
done
