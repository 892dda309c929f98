#!/bin/bash
# round 2, call U: ncu --set full of the op-API upfirdn2d kernels (blur, down-2, up-2; fp32 and bf16)
mkdir -p gpurun_out/r02u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upfirdn2d -o /tmp/ops python tools/probes/op_profile.py > gpurun_out/r02u/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02u/ncu.log
python tools/ncu_summary.py /tmp/ops.ncu-rep > gpurun_out/r02u/ops_summary.json
python tools/ncu_hot.py /tmp/ops.ncu-rep > gpurun_out/r02u/ops_hot.txt 2>&1
ncu -i /tmp/ops.ncu-rep --page details --csv > gpurun_out/r02u/ops_details.csv 2>/dev/null
ls -la /tmp/ops.ncu-rep; cp /tmp/ops.ncu-rep gpurun_out/r02u/ 2>/dev/null
du -sh gpurun_out/r02u
