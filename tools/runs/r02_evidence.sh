#!/bin/bash
# round 2 evidence: ncu --set full of one forward at both benchmark configs (summaries computed on the box, the reports
# themselves stay there: gpurun brings back at most 64 MiB), the launch list of the bench command, the op sweep
mkdir -p gpurun_out/r02ev
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  timeout 900 ncu --profile-from-start off --set full --clock-control none \
      -k regex:'modconv_|upfir_tc|smooth_up|rgb_combine' -o /tmp/step$1 python tools/profile_step.py --size $1 --batch $2 > gpurun_out/r02ev/ncu$1.log 2>&1
  echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/step$1.ncu-rep > gpurun_out/r02ev/ncu_full_r02_v1_$1_b$2_summary.json
  ls -la /tmp/step$1.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ev/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02ev/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 1500 python tools/opbench.py > gpurun_out/r02ev/opbench_r02.jsonl 2> gpurun_out/r02ev/opbench.err
echo "opbench rc=$? lines=$(wc -l < gpurun_out/r02ev/opbench_r02.jsonl)"; tail -3 gpurun_out/r02ev/opbench.err
du -sh gpurun_out/r02ev
