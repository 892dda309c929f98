#!/bin/bash
# round 2, call F: relaxed accumulator hand-back on the cta_group::2 kernels -- parity, then timing with / without the merged polyphase walk
mkdir -p gpurun_out/r02f
python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py tests/test_tc_conv_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02f/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02f/pytest.log
for v in 1 0 2; do
SG2_POLY4=$v timeout 300 python bench.py --no-cpu-baseline --no-extra --profile-out gpurun_out/r02f/k256_poly$v.json > gpurun_out/r02f/bench_poly$v.json 2> gpurun_out/r02f/bench_poly$v.err
done
python tools/kcmp.py gpurun_out/r02f/k256_poly0.json gpurun_out/r02f/k256_poly1.json gpurun_out/r02f/k256_poly2.json
timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02f/k1024.json > gpurun_out/r02f/bench1024.json 2> gpurun_out/r02f/bench1024.err
python tools/kcmp.py gpurun_out/r02f/k1024.json | tail -18
python -c "
import json
for f in ('bench_poly1','bench1024'):
    d=json.load(open('gpurun_out/r02f/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_sum_ms'])"
