#!/bin/bash
# round 2, call Q: final image from the last conv's epilogue with the skip gathered before the accumulator wait
mkdir -p gpurun_out/r02q
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -p no:cacheprovider -x -k "1024" > gpurun_out/r02q/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02q/pytest.log | cut -c1-200
for v in 1 0 1 0; do
SG2_DXS_IMAGE=$v timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02q/k1024_img$v.json > gpurun_out/r02q/b1024_img$v.json 2> gpurun_out/r02q/b1024_img$v.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02q/b1024_img$v.json') if l.startswith('{')][-1]); print($v, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_sum_ms'])"
done
python tools/kcmp.py gpurun_out/r02q/k1024_img0.json gpurun_out/r02q/k1024_img1.json | tail -8
