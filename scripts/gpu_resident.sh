#!/bin/bash
# Resident-kernel visit: its parity tests first (fail fast), smoke, then the configs through the public API.
mkdir -p gpurun_out
echo "== resident tests"; timeout 900 python -m pytest tests/test_gpu_resident.py -q -x --timeout 600 > gpurun_out/pytest_resident.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_resident.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== configs"; timeout 600 python scripts/gpu_configs.py > gpurun_out/configs_res.json 2> gpurun_out/configs_res.err; echo "configs rc=$?"; grep -v Done gpurun_out/configs_res.json | head -50; tail -5 gpurun_out/configs_res.err
