#!/bin/bash
# fp64 (reference precision) path: timing, launch list and one full ncu capture of k_step_lean.  usage: bash scripts/gpu_fp64.sh <tag>
TAG=${1:-a}
mkdir -p gpurun_out
python scripts/prof_fp64.py 4096 24 > gpurun_out/fp64_${TAG}.txt 2>&1; cat gpurun_out/fp64_${TAG}.txt | tail -2
python scripts/prof_fp64.py 4096 24 fp32 '{"rep_gain_C":0.3,"delta_R_D":0.7}' >> gpurun_out/fp64_${TAG}.txt 2>&1; tail -1 gpurun_out/fp64_${TAG}.txt
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fp64_${TAG}.csv python scripts/prof_fp64.py 4096 4 > gpurun_out/ncu_list_fp64_${TAG}.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_step_lean -s 3 -c 1 -o gpurun_out/prof_fp64_${TAG} -f python scripts/prof_fp64.py 4096 4 > gpurun_out/ncu_fp64_${TAG}.log 2>&1; echo "ncu rc=$?"
fi
