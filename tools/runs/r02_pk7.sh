#!/bin/bash
# fp32 down-2 with vector rows in the streaming kernel: op tests + reference-extension tests, then the A/B on the sweep
mkdir -p gpurun_out/pk7
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ref_ext_gpu.py -m gpu -x -q > gpurun_out/pk7/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pk7/pytest.log
for v in 0 1; do
  SG2_UPFIRDN_VEC=$v timeout 600 python tools/opbench.py --no-ref --res 64 128 256 512 1024 --dtypes float32 2>/dev/null | grep down2 > gpurun_out/pk7/down2_vec$v.jsonl
done
python - <<'PY'
import json,statistics
for v in (0,1):
    rows=[json.loads(l) for l in open(f'gpurun_out/pk7/down2_vec{v}.jsonl')]
    print('vec',v,'median',statistics.median(r['frac_of_hbm_peak'] for r in rows),'min',min(r['frac_of_hbm_peak'] for r in rows),'max',max(r['frac_of_hbm_peak'] for r in rows))
PY
