#!/bin/bash
# round 2, call Y: do the epilogues' noise loads sit on the critical path of the GEMM kernels?  (knock-out bit 32: no noise loads)
mkdir -p gpurun_out/r02y
nvidia-smi -q -d POWER | grep -i "power limit\|power draw\|cap" | head -12
export SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200_ko.so
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  names=()
  for d in 0 32; do
    SG2_BENCH_NO_PARITY=1 SG2_GEMM_DBG=$d timeout 200 python bench.py --size $1 --batch $2 --no-cpu-baseline --no-extra --steps 5 --warmup 3 \
        --profile-out gpurun_out/r02y/g$1_$d.json > gpurun_out/r02y/g$1_$d.log 2>&1 || echo "FAILED $d"
    names+=(gpurun_out/r02y/g$1_$d.json)
  done
  python tools/kcmp.py --min-ms 0.2 "${names[@]}" | cut -c1-120
done
