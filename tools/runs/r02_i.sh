#!/bin/bash
# round 2, call I: the training engine -- parity, then the decoder forward + backward step and the fine-tuning step
mkdir -p gpurun_out/r02i
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s -p no:cacheprovider -x > gpurun_out/r02i/pytest.log 2>&1
echo "pytest rc=$?"; grep "parity\]" gpurun_out/r02i/pytest.log; tail -25 gpurun_out/r02i/pytest.log | cut -c1-220
