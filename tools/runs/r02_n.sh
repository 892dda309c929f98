#!/bin/bash
# round 2, call N: the ADA decoder on the engine -- parity (+ the rest of the ADA / engine tests), then speed
mkdir -p gpurun_out/r02n
timeout 900 python -m pytest tests/test_ada_gpu.py tests/test_engine_gpu.py -m gpu -q -s -p no:cacheprovider -x > gpurun_out/r02n/pytest.log 2>&1
echo "pytest rc=$?"; grep "parity\] ADA engine" gpurun_out/r02n/pytest.log; tail -15 gpurun_out/r02n/pytest.log | cut -c1-220
timeout 300 python - > gpurun_out/r02n/ada_speed.json 2> gpurun_out/r02n/ada_speed.err <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import sg2_b200 as sg2
dev = "cuda:0"
out = {}
for name, make in (("ada", lambda: sg2.stylegan2_ada.Generator(512, 512, 8, 256, 3)), ("rosinality", lambda: sg2.Generator(256, 512, 8))):
    torch.manual_seed(0)
    G = make().to(dev).eval()
    z = torch.randn(64, 512, device=dev)
    for mode in ("bf16",) + (("bf16-noengine",) if name == "ada" else ()):
        G.precision = "bf16"
        os.environ["SG2_B200_ADA_ENGINE"] = "0" if mode.endswith("noengine") else "1"
        with torch.no_grad():
            for _ in range(5):
                G([z], randomize_noise=False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                G([z], randomize_noise=False)
            b.record()
            torch.cuda.synchronize()
        out[f"{name}:{mode}"] = {"ms_per_forward_b64_256": round(a.elapsed_time(b) / 20, 3), "img_per_s": round(64 * 20 / a.elapsed_time(b) * 1e3, 1)}
print(json.dumps(out))
PY
cat gpurun_out/r02n/ada_speed.json; tail -3 gpurun_out/r02n/ada_speed.err
