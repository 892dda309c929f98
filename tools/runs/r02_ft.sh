#!/bin/bash
# round 2: the fine-tuning step (BASELINE configs[2]) on N = $1 ranks over NCCL
N=${1:-2}
mkdir -p gpurun_out/r02ft
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 600 $TR --master-port 29511 tools/finetune_bench.py --batch 8 --steps 20 > gpurun_out/r02ft/finetune_n$N.json 2> gpurun_out/r02ft/finetune_n$N.err
echo "finetune N=$N rc=$?"; grep '^{' gpurun_out/r02ft/finetune_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['decoder_ms_per_step'], d['exchange'])"
grep -v "Warning\|OMP\|\*\*\*" gpurun_out/r02ft/finetune_n$N.err | tail -3
