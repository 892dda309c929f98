#!/bin/bash
# round 2, call AC: separable instantiation of the streaming upfirdn2d kernel: op tests, reference-extension comparison, quick op sweep
mkdir -p gpurun_out/r02ac
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ref_ext_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02ac/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02ac/pytest.log | cut -c1-200
timeout 900 python tools/opbench.py --quick --no-ref > gpurun_out/r02ac/opbench_quick.jsonl 2> gpurun_out/r02ac/opbench.err
echo "opbench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02ac/opbench_quick.jsonl'):
    r=json.loads(l)
    if r['op'].startswith('upfirdn2d') and r['res']>=64 and r['C']==64: print(r['op'], r['dtype'], r['res'], r['frac_of_hbm_peak'])
PY
