#!/bin/bash
# round 2, call C: knock-out analysis of the 1024^2 tail (variant library with -DSG2_KNOCKOUT=1; results wrong on purpose)
mkdir -p gpurun_out/r02c
export SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200_ko.so
for d in 0 1 2 8 16 10 17 27; do
  SG2_GEMM_DBG=$d SG2_BENCH_NO_PARITY=1 timeout 200 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --steps 5 \
      --profile-out gpurun_out/r02c/k_dbg$d.json > gpurun_out/r02c/b_dbg$d.json 2> gpurun_out/r02c/b_dbg$d.err || echo "FAILED $d"
done
python tools/kcmp.py --kind gemm gpurun_out/r02c/k_dbg0.json gpurun_out/r02c/k_dbg1.json gpurun_out/r02c/k_dbg2.json gpurun_out/r02c/k_dbg8.json gpurun_out/r02c/k_dbg16.json gpurun_out/r02c/k_dbg10.json gpurun_out/r02c/k_dbg17.json gpurun_out/r02c/k_dbg27.json
tail -2 gpurun_out/r02c/b_dbg2.err
