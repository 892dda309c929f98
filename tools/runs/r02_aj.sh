#!/bin/bash
# round 2, call AJ: big-tile exact conv with table-driven weight staging: parity + timing
mkdir -p gpurun_out/r02aj
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py tests/test_ada_gpu.py tests/test_tc_conv_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02aj/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02aj/pytest.log | cut -c1-200
for cfg in "256 32" "1024 4"; do
  set -- $cfg
  timeout 300 python bench.py --precision exact --size $1 --batch $2 --no-cpu-baseline --no-extra --steps 5 --warmup 3 > gpurun_out/r02aj/exact_$1.json 2> gpurun_out/r02aj/exact_$1.err
  echo "size=$1 B=$2: $(grep -o '"value": [0-9.]*' gpurun_out/r02aj/exact_$1.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02aj/exact_$1.json | head -1)"
done
