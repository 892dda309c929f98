#!/bin/bash
# round 2, call R: ADA engine incl. the zero-padded 16-channel block at 1024^2; training engine timing after the grid change
mkdir -p gpurun_out/r02r
timeout 900 python -m pytest tests/test_ada_gpu.py -m gpu -q -s -p no:cacheprovider -x -k "engine" > gpurun_out/r02r/pytest.log 2>&1
echo "pytest rc=$?"; grep "parity\] ADA engine" gpurun_out/r02r/pytest.log | sort -u; tail -3 gpurun_out/r02r/pytest.log | cut -c1-200
for b in 8 32; do timeout 300 python tools/train_step_bench.py --batch $b --iters 10 2>/dev/null | grep '"bf16"'; done
timeout 300 python - <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import sg2_b200 as sg2
dev = "cuda:0"
torch.manual_seed(0)
G = sg2.stylegan2_ada.Generator(512, 512, 8, 1024, 3).to(dev).eval()
z = torch.randn(4, 512, device=dev)
out = {}
for mode in ("bf16", "bf16-noengine"):
    G.precision = "bf16"
    os.environ["SG2_B200_ADA_ENGINE"] = "0" if mode.endswith("noengine") else "1"
    with torch.no_grad():
        for _ in range(3):
            G([z], randomize_noise=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            G([z], randomize_noise=False)
        b.record()
        torch.cuda.synchronize()
    out["ada1024_b4:" + mode] = round(a.elapsed_time(b) / 10, 3)
print(json.dumps(out))
PY
