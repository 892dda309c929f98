#!/bin/bash
# round 2, call J: speed of the training engine -- decoder forward + backward, the fine-tuning step, a torch.profiler table
mkdir -p gpurun_out/r02j
for b in 8 32; do
  timeout 300 python tools/train_step_bench.py --batch $b --iters 10 >> gpurun_out/r02j/train_step.jsonl 2>> gpurun_out/r02j/train_step.err
done
cat gpurun_out/r02j/train_step.jsonl
timeout 600 python tools/finetune_bench.py --batch 8 --steps 10 > gpurun_out/r02j/finetune_n1.json 2> gpurun_out/r02j/finetune_n1.err
cat gpurun_out/r02j/finetune_n1.json; tail -2 gpurun_out/r02j/finetune_n1.err
timeout 300 python tools/train_step_profile.py > gpurun_out/r02j/train_step_profile.txt 2>&1
head -40 gpurun_out/r02j/train_step_profile.txt | cut -c1-180
