#!/bin/bash
# round 2, call W: knock-out table of the FIR kernel as shipped (loads 1 / stores 2 / TMEM load + math 4 / one MMA 8 and combinations)
mkdir -p gpurun_out/r02w
export SG2_B200_LIB=$PWD/stylegan-for-facerec_b200/csrc/libsg2_b200_ko.so
names=()
for d in 0 16 17 18 20 31; do
  SG2_BENCH_NO_PARITY=1 SG2_FIR_DBG=$d timeout 200 python bench.py --size 256 --batch 64 --no-cpu-baseline --no-extra --steps 5 --warmup 3 \
      --profile-out gpurun_out/r02w/fir$d.json > gpurun_out/r02w/fir$d.log 2>&1 || echo "FAILED $d"
  names+=(gpurun_out/r02w/fir$d.json)
done
python tools/kcmp.py --kind upfir "${names[@]}" | cut -c1-180
