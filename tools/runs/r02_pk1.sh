#!/bin/bash
# packed-math upfirdn2d kernel: bit-exactness vs the streaming kernel, timing A/B, memcheck on a few shapes
mkdir -p gpurun_out/pk1
timeout 600 python tools/probes/pk_check.py --perf > gpurun_out/pk1/check.txt 2> gpurun_out/pk1/check.err
echo "check rc=$?"
tail -25 gpurun_out/pk1/check.txt
tail -5 gpurun_out/pk1/check.err
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/pk1/memcheck.log python tools/probes/pk_check.py --quick > gpurun_out/pk1/check_quick.txt 2>&1
echo "memcheck rc=$?"
tail -3 gpurun_out/pk1/check_quick.txt
grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/pk1/memcheck.log | head -8
