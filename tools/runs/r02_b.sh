#!/bin/bash
# round 2, call B: fused up-sampling conv (1024^2 tail) -- parity, then per-kernel times A/B against the polyphase + FIR plan
mkdir -p gpurun_out/r02b
python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -s -p no:cacheprovider -x > gpurun_out/r02b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b/pytest.log
tail -4 gpurun_out/r02b/pytest.log; grep "parity\]" gpurun_out/r02b/pytest.log
for v in fused:1 polyphase:0 fused_nores:1; do
  name=${v%%:*}; f=${v##*:}
  extra=""; [ $name = fused_nores ] && extra="SG2_GEMM_RES2=0"
  env SG2_UPFUSED=$f $extra timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --steps 10 \
      --profile-out gpurun_out/r02b/k1024_$name.json > gpurun_out/r02b/bench1024_$name.json 2> gpurun_out/r02b/bench1024_$name.err
  echo "$name rc=$?"; tail -2 gpurun_out/r02b/bench1024_$name.err
done
python tools/kcmp.py gpurun_out/r02b/k1024_polyphase.json | tail -14
python tools/kcmp.py gpurun_out/r02b/k1024_fused.json | tail -12
python tools/kcmp.py gpurun_out/r02b/k1024_fused_nores.json | tail -12
