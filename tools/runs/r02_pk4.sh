#!/bin/bash
# knock-out table of the packed-math upfirdn2d kernel (variant libraries -DPK_KO=bits: 1 no FMAs, 2 no stores, 4 no copies, 8 no unpack switch)
mkdir -p gpurun_out/pk4
for v in "" _pk1 _pk2 _pk4 _pk8 _pk7; do
  echo "== variant '$v'" | tee -a gpurun_out/pk4/ko.txt
  SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200$v.so timeout 200 python tools/probes/pk_check.py --perf-only 2>&1 | grep -v "^checked" | tee -a gpurun_out/pk4/ko.txt
done
