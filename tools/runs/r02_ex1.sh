#!/bin/bash
# exact fp32 route after the packed-FMA inner loops: parity tests of the exact path, then its step time (256^2 B=32, 1024^2 B=4)
mkdir -p gpurun_out/ex1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q > gpurun_out/ex1/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/ex1/pytest.log
timeout 300 python bench.py --precision exact --no-extra --no-cpu-baseline --size 256 --batch 32 --steps 5 --warmup 3 > gpurun_out/ex1/exact256.json 2> gpurun_out/ex1/exact256.err; echo rc=$?
timeout 300 python bench.py --precision exact --no-extra --no-cpu-baseline --size 1024 --batch 4 --steps 5 --warmup 3 > gpurun_out/ex1/exact1024.json 2> gpurun_out/ex1/exact1024.err; echo rc=$?
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": 1, "steps": 5, "warmup": 3, "ms_per_step": [0-9.]*' gpurun_out/ex1/exact256.json gpurun_out/ex1/exact1024.json
