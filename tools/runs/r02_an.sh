#!/bin/bash
# round 2, call AN: rgb_combine templated on the pooling factor (the plain launch must be as fast as before the face_pool fusion)
mkdir -p gpurun_out/r02an
timeout 900 python -m pytest tests/test_psp_io_gpu.py tests/test_engine_gpu.py tests/test_ada_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02an/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02an/pytest.log | cut -c1-200
for cfg in "1024 32" "256 64"; do
  set -- $cfg
  timeout 300 python bench.py --size $1 --batch $2 --no-cpu-baseline --no-extra --profile-out gpurun_out/r02an/k$1.json > gpurun_out/r02an/b$1.json 2> gpurun_out/r02an/b$1.err || echo FAILED
  python tools/kcmp.py --min-ms 0.02 gpurun_out/r02an/k$1.json | grep "rgb_c\|total"
done
python - <<'PY'
import importlib, sys, os, torch
sys.path.insert(0, os.getcwd())
sg2 = importlib.import_module("stylegan-for-facerec_b200")
io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
G = sg2.Generator(1024, 512, 8).to("cuda:0").eval(); G.precision = "bf16"
z = torch.randn(32, 512, device="cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
with torch.no_grad():
    t1 = t(lambda: io.face_pool(G([z], randomize_noise=False)[0], 256))
    t2 = t(lambda: io.decode_pooled(G, [z], 256, randomize_noise=False)[0])
print(f"1024^2 B=32 decoder + face_pool: {t1:.3f} ms; decode_pooled (fused): {t2:.3f} ms")
PY
