#!/bin/bash
# round 2, call V: FIR with two epilogue groups (SG2_FIR_GROUPS=2, Toeplitz matrix in TMEM) vs the shipped kernel: parity + per-kernel times
mkdir -p gpurun_out/r02v
L=$PWD/stylegan-for-facerec_b200/csrc
SG2_B200_LIB=$L/libsg2_b200_fg2.so timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02v/pytest_fg2.log 2>&1
echo "pytest fg2 rc=$?"; tail -2 gpurun_out/r02v/pytest_fg2.log | cut -c1-200
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  for v in base fg2 fa1; do
    lib=$L/libsg2_b200.so; [ $v != base ] && lib=$L/libsg2_b200_$v.so
    SG2_B200_LIB=$lib timeout 300 python bench.py --size $1 --batch $2 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02v/k$1_$v.json > gpurun_out/r02v/b$1_$v.json 2> gpurun_out/r02v/b$1_$v.err || echo "FAILED $v $1"
  done
  python tools/kcmp.py --kind upfir gpurun_out/r02v/k$1_base.json gpurun_out/r02v/k$1_fg2.json gpurun_out/r02v/k$1_fa1.json
done
grep -o '"value": [0-9.]*' gpurun_out/r02v/b*_*.json | head -12
