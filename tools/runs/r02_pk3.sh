#!/bin/bash
mkdir -p gpurun_out/pk3
timeout 600 python tools/probes/pk_check.py --perf > gpurun_out/pk3/check.txt 2> gpurun_out/pk3/check.err
echo "check rc=$?"
grep -v float16 gpurun_out/pk3/check.txt | tail -14
tail -5 gpurun_out/pk3/check.err
