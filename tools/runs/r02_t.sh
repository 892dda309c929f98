#!/bin/bash
# round 2, call T: full -m gpu suite (the command the driver runs at round end) on the current tree, then the ncu capture of the op kernels
mkdir -p gpurun_out/r02t gpurun_out/r02u
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02t/pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r02t/pytest.log | cut -c1-200
bash tools/runs/r02_u.sh
