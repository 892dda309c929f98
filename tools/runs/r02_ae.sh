#!/bin/bash
# round 2, call AE: dx-stacked kernel with the incremental tile walk / noise one iteration ahead / no packing on the last layer
mkdir -p gpurun_out/r02ae
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02ae/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02ae/pytest.log | cut -c1-200
timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02ae/k1024.json > gpurun_out/r02ae/b1024.json 2> gpurun_out/r02ae/b1024.err || echo "FAILED"
python tools/kcmp.py --min-ms 0.2 gpurun_out/r02ae/k1024.json
