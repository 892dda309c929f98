#!/bin/bash
# round 2, call S: packed f32x2 epilogue of the dx-stacked kernel -- parity + timing
mkdir -p gpurun_out/r02s
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -p no:cacheprovider -x -k "1024" > gpurun_out/r02s/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02s/pytest.log | cut -c1-200
for v in 1 0; do
SG2_DXS=$v timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02s/k1024_dxs$v.json > gpurun_out/r02s/b1024_dxs$v.json 2> gpurun_out/r02s/b1024_dxs$v.err
done
python tools/kcmp.py gpurun_out/r02s/k1024_dxs0.json gpurun_out/r02s/k1024_dxs1.json | tail -9
