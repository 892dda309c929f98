#!/bin/bash
# round 2, final evidence of the third session: full -m gpu suite, smoke, the bench line (both arms), the launch list of the bench
# command, the op sweep (BASELINE configs[4]) with clocks and the reference's own kernels, ncu --set full of the op kernels
mkdir -p gpurun_out/r02f2
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02f2/pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02f2/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f2/smoke.log 2>&1; echo "smoke rc=$?"
S=$(date +%s)
timeout 900 python bench.py --profile-out gpurun_out/r02f2/k256.json > gpurun_out/r02f2/bench.json 2> gpurun_out/r02f2/bench.err
echo "bench rc=$? $(( $(date +%s) - S )) s"
timeout 600 python bench.py --impl reference > gpurun_out/r02f2/bench_ref.json 2> gpurun_out/r02f2/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f2/launches_r02_v3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02f2/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 1500 python tools/opbench.py > gpurun_out/r02f2/opbench_r02_v3.jsonl 2> gpurun_out/r02f2/opbench.err
echo "opbench rc=$? lines=$(wc -l < gpurun_out/r02f2/opbench_r02_v3.jsonl)"
timeout 600 ncu --set full --clock-control none -k regex:upfirdn2d -o /tmp/ops python tools/probes/op_profile.py > gpurun_out/r02f2/ncu_ops.log 2>&1
echo "ncu ops rc=$?"
python tools/ncu_summary.py /tmp/ops.ncu-rep > gpurun_out/r02f2/ncu_full_r02_v3_upfirdn2d_summary.json
du -sh gpurun_out/r02f2
