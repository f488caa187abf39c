#!/bin/bash
mkdir -p gpurun_out
echo "== resident tests"; timeout 900 python -m pytest tests/test_gpu_resident.py -q -x --timeout 600 > gpurun_out/pytest_resident.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_resident.log
echo "== res_bench"; timeout 600 python scripts/res_bench.py > gpurun_out/res_bench.log 2>&1; echo "rc=$?"; grep -v "^{" gpurun_out/res_bench.log | head -40
echo "== parity tests touched by the path pinning"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 -k "chunking or batched or fast_path or philox" > gpurun_out/pytest_parity_sub.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_parity_sub.log
RCMD="python bench.py --workload sweep --steps 1 --warmup 1 --inner 300"
$RCMD > gpurun_out/plain_res_r01t.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_res_r01t.csv $RCMD > gpurun_out/ncu_list_res_r01t.log 2>&1; echo "list rc=$?"
