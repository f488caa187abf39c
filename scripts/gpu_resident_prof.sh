#!/bin/bash
# ncu --set full of the resident kernel on one small case (argument: res_bench case name)
mkdir -p gpurun_out
CASE=${1:-c1_L64}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 1 -s 1 -o gpurun_out/prof_res_$CASE -f python scripts/res_bench.py $CASE > gpurun_out/ncu_res_$CASE.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_res_$CASE.log
