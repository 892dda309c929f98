#!/bin/bash
# round 2, call AB: ncu --set full with source of the 1024^2 tail GEMM launches (128->64 up4, 64->64, 64->32 up4, 32->32 dx-stacked)
mkdir -p gpurun_out/r02ab
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'modconv_' --launch-skip 12 --launch-count 5 -o /tmp/tail python tools/profile_step.py --size 1024 --batch 32 > gpurun_out/r02ab/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02ab/ncu.log
python tools/ncu_summary.py /tmp/tail.ncu-rep > gpurun_out/r02ab/tail_summary.json
python tools/ncu_hot.py /tmp/tail.ncu-rep > gpurun_out/r02ab/tail_hot.txt 2>&1
ncu -i /tmp/tail.ncu-rep --page details --csv > gpurun_out/r02ab/tail_details.csv 2>/dev/null
ls -la /tmp/tail.ncu-rep; cp /tmp/tail.ncu-rep gpurun_out/r02ab/ 2>/dev/null; du -sh gpurun_out/r02ab
