#!/bin/bash
mkdir -p gpurun_out
echo "== res_bench"; timeout 600 python scripts/res_bench.py > gpurun_out/res_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/res_bench.log | head -40
echo "== ncu resident"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident -c 1 -s 1 -o gpurun_out/prof_res -f python scripts/res_bench.py c2_L200_act_m2 > gpurun_out/ncu_res.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/ncu_res.log
