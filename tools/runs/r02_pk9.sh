#!/bin/bash
# ncu --set full of the streaming kernels after the packed up-2 / vector-row changes
mkdir -p gpurun_out/pk9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:upfirdn2d_stream -o /tmp/us python tools/probes/op_profile.py > gpurun_out/pk9/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/pk9/ncu.log
python tools/ncu_summary.py /tmp/us.ncu-rep > gpurun_out/pk9/us_summary.json
ncu -i /tmp/us.ncu-rep --page raw --csv > gpurun_out/pk9/us_raw.csv 2>/dev/null
