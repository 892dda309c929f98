#!/bin/bash
# round 2, call AG: E-pass of the training engine with a cp.async input ring: gradient parity + timings
mkdir -p gpurun_out/r02ag
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r02ag/pytest.log 2>&1
echo "pytest rc=$?"; grep -i "rel\|parity" gpurun_out/r02ag/pytest.log | sort -u | head -12 | cut -c1-200; tail -2 gpurun_out/r02ag/pytest.log | cut -c1-200
for b in 8 32; do timeout 300 python tools/train_step_bench.py --batch $b --iters 10 2>/dev/null | grep '"bf16"' | tee -a gpurun_out/r02ag/train_step.jsonl; done
timeout 300 python tools/train_step_profile.py > gpurun_out/r02ag/train_step_profile.txt 2>&1; grep "epass\|firT\|gemm2_kernel\|Self CUDA time" gpurun_out/r02ag/train_step_profile.txt | cut -c1-70,140-215
timeout 600 python bench.py --workload finetune > gpurun_out/r02ag/finetune_n1.json 2> gpurun_out/r02ag/finetune_n1.err; echo "finetune rc=$?"; cut -c1-260 gpurun_out/r02ag/finetune_n1.json
