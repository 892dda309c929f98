#!/bin/bash
# round 2, call AH: register-blocked big tiles in the exact fp32 modulated conv: golden parity + exact-path timings (old / new)
mkdir -p gpurun_out/r02ah
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py tests/test_ada_gpu.py tests/test_tc_conv_gpu.py tests/test_bench_configs_gpu.py tests/test_psp_io_gpu.py tests/test_engine_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02ah/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02ah/pytest.log | cut -c1-200
for v in 0 1; do
  for cfg in "256 32" "1024 4"; do
    set -- $cfg
    SG2_MODCONV_BIG=$v timeout 300 python bench.py --precision exact --size $1 --batch $2 --no-cpu-baseline --no-extra --steps 5 --warmup 3 > gpurun_out/r02ah/exact_$1_big$v.json 2> gpurun_out/r02ah/exact_$1_big$v.err
    echo "big=$v size=$1 B=$2: $(grep -o '"value": [0-9.]*' gpurun_out/r02ah/exact_$1_big$v.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02ah/exact_$1_big$v.json | head -1)"
  done
done
for v in 0 1; do SG2_MODCONV_BIG=$v timeout 300 python tools/train_step_bench.py --batch 8 --iters 5 2>/dev/null | grep '"exact"' | cut -c1-200; done
