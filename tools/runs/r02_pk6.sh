#!/bin/bash
# full GPU test suite with the packed upfirdn2d kernel in place, then the op sweep (BASELINE configs[4]) with clocks and the reference kernels
mkdir -p gpurun_out/pk6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pk6/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pk6/pytest.log
timeout 900 python tools/opbench.py > gpurun_out/pk6/opbench.jsonl 2> gpurun_out/pk6/opbench.err
echo "opbench rc=$?"; wc -l gpurun_out/pk6/opbench.jsonl; tail -2 gpurun_out/pk6/opbench.err
