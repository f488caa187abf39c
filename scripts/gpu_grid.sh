#!/bin/bash
mkdir -p gpurun_out
echo "== all resident tests"; timeout 600 python -m pytest tests/test_gpu_resident.py -q -x --timeout 300 > gpurun_out/pytest_resident.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_resident.log
echo "skip bench"
echo "== whole gpu suite"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
