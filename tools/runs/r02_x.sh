#!/bin/bash
# round 2, call X: FIR kernel with TMA-fed epilogue inputs (noise tile / bias / style ring), Toeplitz matrix in TMEM, double-buffered staging
mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py tests/test_train_engine_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02x/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02x/pytest.log | cut -c1-200
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  timeout 300 python bench.py --size $1 --batch $2 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02x/k$1.json > gpurun_out/r02x/b$1.json 2> gpurun_out/r02x/b$1.err || echo "FAILED $1"
  python tools/kcmp.py --kind upfir gpurun_out/r02v/k$1_base.json gpurun_out/r02x/k$1.json 2>/dev/null || python tools/kcmp.py --kind upfir gpurun_out/r02x/k$1.json
done
grep -o '"value": [0-9.]*' gpurun_out/r02x/b*.json | head -4
