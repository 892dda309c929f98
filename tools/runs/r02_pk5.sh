#!/bin/bash
# A/B: CTAs per SM (5 vs 6), items per warp
mkdir -p gpurun_out/pk5
for v in "" _c6; do
 for it in 2 3 6; do
  echo "== variant '$v' items/warp $it" | tee -a gpurun_out/pk5/ab.txt
  SG2_PK_ITEMS=$it SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200$v.so timeout 200 python tools/probes/pk_check.py --perf-only 2>&1 | grep -v "^checked" | tee -a gpurun_out/pk5/ab.txt
 done
done
