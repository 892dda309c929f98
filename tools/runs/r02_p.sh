#!/bin/bash
# round 2, call P: final image written by the last conv's epilogue (1024^2) -- parity, then timing
mkdir -p gpurun_out/r02p
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_configs_gpu.py tests/test_train_engine_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02p/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02p/pytest.log | cut -c1-200
for v in 1 0; do
SG2_DXS_IMAGE=$v timeout 300 python bench.py --size 1024 --batch 32 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02p/k1024_img$v.json > gpurun_out/r02p/b1024_img$v.json 2> gpurun_out/r02p/b1024_img$v.err
done
python tools/kcmp.py gpurun_out/r02p/k1024_img0.json gpurun_out/r02p/k1024_img1.json | tail -12
python -c "
import json
for v in (0,1):
    d=json.loads([l for l in open('gpurun_out/r02p/b1024_img%d.json'%v) if l.startswith('{')][-1]); print(v, d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['rel_max_vs_exact_fp32'])"
