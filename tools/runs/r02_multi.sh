#!/bin/bash
# round 2, multi-GPU call: N = $1 ranks -- the fine-tuning step (NCCL gradient exchange) and the synthesis bench
N=${1:-2}
mkdir -p gpurun_out/r02m
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/finetune_bench.py --batch 8 --steps 10 > gpurun_out/r02m/finetune_n$N.json 2> gpurun_out/r02m/finetune_n$N.err
echo "finetune N=$N rc=$?"; cat gpurun_out/r02m/finetune_n$N.json; grep -v Warning gpurun_out/r02m/finetune_n$N.err | tail -3
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m/bench_n$N.json 2> gpurun_out/r02m/bench_n$N.err
echo "bench N=$N rc=$?"; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02m/bench_n$N.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['ranks'], d['configs']['1024_b32']['value'], d['configs']['1024_b32']['ranks'])"
grep -v Warning gpurun_out/r02m/bench_n$N.err | tail -3
