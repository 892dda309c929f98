#!/bin/bash
# round 2, call H: ncu evidence for the shipped kernels -- full-set captures of one forward at both benchmark configs + the launch list of the bench command
mkdir -p gpurun_out/r02h
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:'modconv_|upfir_tc|rgb_combine' -o gpurun_out/r02h/step$1 python tools/profile_step.py --size $1 --batch $2 > gpurun_out/r02h/ncu$1.log 2>&1
  echo "ncu $1 rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02h/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; ls -la gpurun_out/r02h
