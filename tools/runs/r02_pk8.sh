#!/bin/bash
# up-2 on packed row-pair FMAs in the streaming kernel: op tests + reference-extension tests, then the sweep lines of up-2
mkdir -p gpurun_out/pk8
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ref_ext_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/pk8/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pk8/pytest.log
timeout 600 python tools/opbench.py --no-ref --res 64 128 256 512 2>/dev/null | grep up2 > gpurun_out/pk8/up2.jsonl
python - <<'PY'
import json,statistics,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/pk8/up2.jsonl'):
    r=json.loads(l); d[r['dtype']].append(r['frac_of_hbm_peak'])
for k,v in d.items(): print(k,'median',statistics.median(v),'min',min(v),'max',max(v))
PY
