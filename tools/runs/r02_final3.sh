#!/bin/bash
# last full check of the session: -m gpu suite, smoke, default bench line
mkdir -p gpurun_out/r02f4
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02f4/pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02f4/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f4/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02f4/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02f4/bench.json 2> gpurun_out/r02f4/bench.err; echo "bench rc=$?"
