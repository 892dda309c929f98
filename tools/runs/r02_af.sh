#!/bin/bash
# round 2, call AF: batched small-planes upfirdn2d kernel: op tests (all), then the op sweep at 4^2 .. 32^2 with the kernel on / off
mkdir -p gpurun_out/r02af
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ref_ext_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02af/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02af/pytest.log | cut -c1-200
for v in 0 1 2; do
  SG2_UPFIRDN_PLANES=$v timeout 600 python tools/opbench.py --no-ref --res 4 8 16 32 --dtypes float32 bfloat16 > gpurun_out/r02af/op_planes$v.jsonl 2> gpurun_out/r02af/op_planes$v.err
done
python - <<'PY'
import json
def load(f): return {(r['op'],r['dtype'],r['res'],r['C']):r for r in map(json.loads, open(f))}
A,B,C=[load(f"gpurun_out/r02af/op_planes{v}.jsonl") for v in (0,1,2)]
for k in A:
    if k[0].startswith('upfirdn2d') and k[3] in (64,): print(k, A[k]["frac_of_hbm_peak"], "->", B[k]["frac_of_hbm_peak"], "all:", C[k]["frac_of_hbm_peak"])
PY
