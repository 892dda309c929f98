#!/bin/bash
# round 2, call O: re-test the pair kernel on every layer now that its hand-back is relaxed (round 1 measured it slower on the narrow ones)
mkdir -p gpurun_out/r02o
run() { name=$1; shift; env "$@" timeout 300 python bench.py --size ${SIZE} --batch ${BATCH} --no-cpu-baseline --no-extra --steps 10 \
   --profile-out gpurun_out/r02o/k${SIZE}_$name.json > gpurun_out/r02o/b${SIZE}_$name.json 2> gpurun_out/r02o/b${SIZE}_$name.err || echo FAILED $name; }
SIZE=1024 BATCH=32
run base X=0; run all2sm SG2_GEMM_2SM=2; run all2sm_res SG2_GEMM_2SM=2 SG2_GEMM_RES2=1
python tools/kcmp.py gpurun_out/r02o/k1024_base.json gpurun_out/r02o/k1024_all2sm.json gpurun_out/r02o/k1024_all2sm_res.json
SIZE=256 BATCH=64
run base X=0; run all2sm SG2_GEMM_2SM=2; run all2sm_res SG2_GEMM_2SM=2 SG2_GEMM_RES2=1
python tools/kcmp.py gpurun_out/r02o/k256_base.json gpurun_out/r02o/k256_all2sm.json gpurun_out/r02o/k256_all2sm_res.json | tail -14
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s -p no:cacheprovider -k large > gpurun_out/r02o/pytest_large.log 2>&1
echo "pytest rc=$?"; grep "parity\]" gpurun_out/r02o/pytest_large.log | sort -u; tail -3 gpurun_out/r02o/pytest_large.log | cut -c1-200
