#!/bin/bash
# round 2, call AM: with the FIR kernel 1.5x faster, is the unfused plan (polyphase GEMM + FIR) of the narrow octaves competitive again?
mkdir -p gpurun_out/r02am
names=()
for v in "fused X=0" "unfused SG2_UPFUSED=0" "unfused_p4 SG2_UPFUSED=0 SG2_POLY4=2"; do
  set -- $v; name=$1; shift
  env "$@" timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --steps 10 --warmup 3 --profile-out gpurun_out/r02am/k_$name.json > gpurun_out/r02am/b_$name.json 2> gpurun_out/r02am/b_$name.err || echo "FAILED $name"
  names+=(gpurun_out/r02am/k_$name.json)
done
for f in "${names[@]}"; do echo "== $f"; python tools/kcmp.py --min-ms 0.15 $f | grep "L20\|L21\|L23\|L24\|total"; done
