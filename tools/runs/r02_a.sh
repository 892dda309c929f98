#!/bin/bash
# round 2, call A: parity at the benchmarked configs, the bench line with the new fields, the fine-tuning step on one GPU
mkdir -p gpurun_out/r02a
python -m pytest tests/test_bench_configs_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02a/pytest_benchcfg.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a/pytest_benchcfg.log
timeout 900 python bench.py --profile-out gpurun_out/r02a/kernels_256.json > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err
echo "bench rc=$?"
timeout 600 python tools/finetune_bench.py --batch 8 --steps 5 > gpurun_out/r02a/finetune_n1.json 2> gpurun_out/r02a/finetune_n1.err
echo "finetune rc=$?"
tail -5 gpurun_out/r02a/pytest_benchcfg.log
grep "parity\]" gpurun_out/r02a/pytest_benchcfg.log
cat gpurun_out/r02a/bench.json | head -c 6000
tail -3 gpurun_out/r02a/bench.err
cat gpurun_out/r02a/finetune_n1.json; tail -3 gpurun_out/r02a/finetune_n1.err
