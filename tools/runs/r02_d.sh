#!/bin/bash
# round 2, call D: 16 epilogue warps on the cta_group::2 kernel + merged polyphase walk -- parity, then timing
mkdir -p gpurun_out/r02d
python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py tests/test_tc_conv_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02d/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02d/pytest.log
SG2_POLY4=2 python -m pytest tests/test_engine_gpu.py tests/test_tc_conv_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02d/pytest_poly4all.log 2>&1
echo "pytest(poly4 everywhere) rc=$?"; tail -3 gpurun_out/r02d/pytest_poly4all.log
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/r02d/k256.json > gpurun_out/r02d/bench.json 2> gpurun_out/r02d/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r02d/bench.err
SG2_POLY4=0 timeout 300 python bench.py --no-cpu-baseline --no-extra --profile-out gpurun_out/r02d/k256_nopoly.json > gpurun_out/r02d/bench_nopoly.json 2> gpurun_out/r02d/bench_nopoly.err
SG2_POLY4=2 timeout 300 python bench.py --no-cpu-baseline --no-extra --profile-out gpurun_out/r02d/k256_polyall.json > gpurun_out/r02d/bench_polyall.json 2> gpurun_out/r02d/bench_polyall.err
python tools/kcmp.py gpurun_out/r02d/k256_nopoly.json gpurun_out/r02d/k256.json gpurun_out/r02d/k256_polyall.json
python tools/kcmp.py gpurun_out/r02d/k256_1024.json | tail -16
python -c "
import json; d=json.load(open('gpurun_out/r02d/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['configs']['1024_b32']['value'], d['configs']['1024_b32']['ms_per_step'])"
