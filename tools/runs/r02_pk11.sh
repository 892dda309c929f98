#!/bin/bash
mkdir -p gpurun_out/pk11
for v in "" _rb4; do
  echo "== variant '$v'" | tee -a gpurun_out/pk11/rb.txt
  SG2_UPFIRDN_PK_QS=$Q SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200$v.so timeout 200 python tools/probes/pk_check.py --perf-only 2>&1 | grep '"op"' | tee -a gpurun_out/pk11/rb.txt
done
