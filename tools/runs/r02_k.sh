#!/bin/bash
# round 2, call K: training engine after the E-pass / FIR-adjoint rework -- parity, speed, profile
mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s -p no:cacheprovider -x > gpurun_out/r02k/pytest.log 2>&1
echo "pytest rc=$?"; grep "parity\]" gpurun_out/r02k/pytest.log | head -4; tail -3 gpurun_out/r02k/pytest.log | cut -c1-200
for b in 8 32; do
  timeout 300 python tools/train_step_bench.py --batch $b --iters 10 >> gpurun_out/r02k/train_step.jsonl 2>> gpurun_out/r02k/train_step.err
done
grep '"bf16"' gpurun_out/r02k/train_step.jsonl
timeout 300 python tools/train_step_profile.py > gpurun_out/r02k/train_step_profile.txt 2>&1
grep -E "train_|gemm|upfir|demod|addmm|rgb_" gpurun_out/r02k/train_step_profile.txt | cut -c1-70,150-200 | head -16
timeout 600 python tools/finetune_bench.py --batch 8 --steps 10 > gpurun_out/r02k/finetune_n1.json 2> gpurun_out/r02k/finetune_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02k/finetune_n1.json')); print(d['value'], d['ms_per_step'], d['decoder_ms_per_step'])"
