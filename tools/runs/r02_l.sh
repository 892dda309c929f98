#!/bin/bash
# round 2, call L: the whole -m gpu suite + the default bench line
mkdir -p gpurun_out/r02l
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02l/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02l/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --profile-out gpurun_out/r02l/kernels_256.json > gpurun_out/r02l/bench.json 2> gpurun_out/r02l/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r02l/bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02l/bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_sum_ms'], d['clocks'])
c=d['configs']['1024_b32']; print(c['value'], c['ms_per_step'], c['e2e']['value'], c['roofline']['kernel_sum_ms'], c['clocks'])
print(d['gpu_reference'])"
