#!/bin/bash
# round 2, call G: dx-stacked narrow convs + early accumulator hand-back -- parity, then timing
mkdir -p gpurun_out/r02g
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py tests/test_tc_conv_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02g/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02g/pytest.log
for v in 1 0; do
SG2_DXS=$v timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02g/k1024_dxs$v.json > gpurun_out/r02g/bench1024_dxs$v.json 2> gpurun_out/r02g/bench1024_dxs$v.err
tail -2 gpurun_out/r02g/bench1024_dxs$v.err
done
python tools/kcmp.py gpurun_out/r02g/k1024_dxs0.json gpurun_out/r02g/k1024_dxs1.json | tail -24
timeout 300 python bench.py --no-cpu-baseline --no-extra --profile-out gpurun_out/r02g/k256.json > gpurun_out/r02g/bench256.json 2> gpurun_out/r02g/bench256.err
python tools/kcmp.py gpurun_out/r02g/k256.json | tail -14
python -c "
import json
for f in ('bench256','bench1024_dxs1'):
    d=json.load(open('gpurun_out/r02g/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_sum_ms'], d['parity'])"
