#!/bin/bash
# round 2, call Z: GPU timeline of three bench steps (what sits between the engine's kernels)
mkdir -p gpurun_out/r02z
timeout 300 python tools/timeline.py --size 256 --batch 64 --steps 3 > gpurun_out/r02z/timeline256.txt 2> gpurun_out/r02z/timeline256.err
echo "rc=$?"; tail -5 gpurun_out/r02z/timeline256.txt
