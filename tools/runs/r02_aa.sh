#!/bin/bash
# round 2, call AA: noise fetched one tile ahead in the pair kernel (up4 layers) and the dx-stacked kernel: parity + 1024^2 timing
mkdir -p gpurun_out/r02aa
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02aa/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02aa/pytest.log | cut -c1-200
for cfg in "1024 32" "256 64"; do
  set -- $cfg
  timeout 300 python bench.py --size $1 --batch $2 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02aa/k$1.json > gpurun_out/r02aa/b$1.json 2> gpurun_out/r02aa/b$1.err || echo "FAILED $1"
  python tools/kcmp.py --min-ms 0.2 gpurun_out/r02aa/k$1.json
done
grep -o '"value": [0-9.]*' gpurun_out/r02aa/b*.json | head -4
