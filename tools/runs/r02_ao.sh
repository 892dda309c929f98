#!/bin/bash
# round 2, call AO: training engine replayed from CUDA graphs (forward graph + backward graph incl. the dL/dlatent products)
mkdir -p gpurun_out/r02ao
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_engine_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02ao/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02ao/pytest.log | cut -c1-200
for v in 0 1; do
  for b in 8 32; do SG2_B200_TRAIN_GRAPH=$v timeout 300 python tools/train_step_bench.py --batch $b --iters 20 2>/dev/null | grep '"bf16"' | sed "s/^/graph=$v /" | cut -c1-200 | tee -a gpurun_out/r02ao/train_step.jsonl; done
  SG2_B200_TRAIN_GRAPH=$v timeout 600 python bench.py --workload finetune > gpurun_out/r02ao/finetune_g$v.json 2> gpurun_out/r02ao/finetune_g$v.err; echo "finetune graph=$v rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/r02ao/finetune_g$v.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02ao/finetune_g$v.json | head -1)"
done
