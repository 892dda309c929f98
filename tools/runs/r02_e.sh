#!/bin/bash
# round 2, call E: ncu --set full of one 1024^2 forward (GEMM launches of the 256^2..1024^2 octaves) with source
mkdir -p gpurun_out/r02e
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'modconv_gemm' -o gpurun_out/r02e/step1024 python tools/profile_step.py --size 1024 --batch 32 > gpurun_out/r02e/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02e/ncu.log; ls -la gpurun_out/r02e/
