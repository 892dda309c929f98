#!/bin/bash
# knock-out table of the packed kernel, final structure (1 no FMAs, 2 no stores -- compute kept alive, 4 no copies)
mkdir -p gpurun_out/pk12
for v in "" _pk1 _pk2 _pk4; do
  echo "== variant '$v'" | tee -a gpurun_out/pk12/ko.txt
  SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200$v.so timeout 200 python tools/probes/pk_check.py --perf-only 2>&1 | grep '"op"' | tee -a gpurun_out/pk12/ko.txt
done
