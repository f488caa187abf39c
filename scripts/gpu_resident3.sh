#!/bin/bash
mkdir -p gpurun_out
echo "== resident tests"; timeout 900 python -m pytest tests/test_gpu_resident.py -q -x --timeout 600 > gpurun_out/pytest_resident.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_resident.log
echo "== res_bench"; timeout 600 python scripts/res_bench.py > gpurun_out/res_bench.log 2>&1; echo "rc=$?"; grep resident gpurun_out/res_bench.log | grep -v "^{" | head -40
echo "== trace"; timeout 300 python scripts/res_trace.py > gpurun_out/res_trace.log 2>&1; echo "rc=$?"
