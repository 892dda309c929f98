#!/bin/bash
# round 2, end-of-session evidence: full -m gpu suite, smoke, the bench line (both arms), ncu --set full of one forward at both
# benchmark configs (summaries only), the launch list of the bench command, decoder fwd+bwd and the fine-tuning step on one GPU
mkdir -p gpurun_out/r02f1
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02f1/pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02f1/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f1/smoke.log 2>&1; echo "smoke rc=$?"
S=$(date +%s)
timeout 900 python bench.py --profile-out gpurun_out/r02f1/k256.json > gpurun_out/r02f1/bench.json 2> gpurun_out/r02f1/bench.err
echo "bench rc=$? $(( $(date +%s) - S )) s"
timeout 600 python bench.py --impl reference > gpurun_out/r02f1/bench_ref.json 2> gpurun_out/r02f1/bench_ref.err; echo "ref rc=$?"
for cfg in "256 64" "1024 32"; do
  set -- $cfg
  timeout 900 ncu --profile-from-start off --set full --clock-control none \
      -k regex:'modconv_|upfir_tc|smooth_up|rgb_combine' -o /tmp/step$1 python tools/profile_step.py --size $1 --batch $2 > gpurun_out/r02f1/ncu$1.log 2>&1
  echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/step$1.ncu-rep > gpurun_out/r02f1/ncu_full_r02_v2_$1_b$2_summary.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f1/launches_r02_v2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02f1/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
for b in 8 32; do timeout 300 python tools/train_step_bench.py --batch $b --iters 10 2>/dev/null | grep '"bf16"' | tee -a gpurun_out/r02f1/train_step.jsonl; done
timeout 600 python bench.py --workload finetune > gpurun_out/r02f1/finetune_n1.json 2> gpurun_out/r02f1/finetune_n1.err; echo "finetune rc=$?"; cut -c1-300 gpurun_out/r02f1/finetune_n1.json
python - <<'PY'
import torch, time
x = torch.empty(100 << 20, dtype=torch.uint8, device='cuda'); h = torch.empty(100 << 20, dtype=torch.uint8).pin_memory()
for _ in range(3): h.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): h.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print(f"pinned D2H 100 MiB: {dt*1e3:.2f} ms = {(100<<20)/dt/1e9:.1f} GB/s")
PY
du -sh gpurun_out/r02f1
