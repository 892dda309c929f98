#!/bin/bash
# bottleneck analysis: per-kernel CUDA-event tables with parts of the FIR / GEMM kernels knocked out
# (SG2_FIR_DBG / SG2_GEMM_DBG; results are wrong on purpose).  usage: tools/knockout.sh [size] [batch]
SIZE=${1:-256}; BATCH=${2:-64}
mkdir -p gpurun_out/ko
# the knock-out branches are compiled only into a variant build of the library (-DSG2_KNOCKOUT=1), selected by SG2_B200_LIB
KO_LIB=stylegan-for-facerec_b200/csrc/libsg2_b200_ko.so
[ -f $KO_LIB ] || python stylegan-for-facerec_b200/build.py --tag ko -DSG2_KNOCKOUT=1
export SG2_B200_LIB=$PWD/$KO_LIB
run() { # name, env...
  name=$1; shift
  env SG2_BENCH_NO_PARITY=1 "$@" timeout 200 python bench.py --size $SIZE --batch $BATCH --no-cpu-baseline --no-extra --steps 5 --warmup 3 \
      --profile-out gpurun_out/ko/$name.json > gpurun_out/ko/$name.log 2>&1 || echo "FAILED $name"
}
run base X=0
for d in 1 2 4 8 3 7; do run fir$d SG2_FIR_DBG=$d; done
run g1sm SG2_GEMM_2SM=0
for d in 1 2 4 6 8 7 15; do run g1sm_dbg$d SG2_GEMM_2SM=0 SG2_GEMM_DBG=$d; done
python tools/kcmp.py gpurun_out/ko/base.json gpurun_out/ko/fir*.json > gpurun_out/ko/fir_table.txt
python tools/kcmp.py gpurun_out/ko/base.json gpurun_out/ko/g1sm*.json > gpurun_out/ko/gemm_table.txt
