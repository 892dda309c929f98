#!/bin/bash
# ncu --set full of the packed-math upfirdn2d kernel (bf16 blur / down-2, 256^2 planes)
mkdir -p gpurun_out/pk2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upfirdn2d_pk -o /tmp/pk python tools/probes/pk_profile.py > gpurun_out/pk2/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/pk2/ncu.log
python tools/ncu_summary.py /tmp/pk.ncu-rep > gpurun_out/pk2/pk_summary.json
python tools/ncu_hot.py /tmp/pk.ncu-rep > gpurun_out/pk2/pk_hot.txt 2>&1
ncu -i /tmp/pk.ncu-rep --page details --csv > gpurun_out/pk2/pk_details.csv 2>/dev/null
cp /tmp/pk.ncu-rep gpurun_out/pk2/ 2>/dev/null
du -sh gpurun_out/pk2
