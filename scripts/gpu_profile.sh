#!/bin/bash
# ncu evidence for the bench command (short variant): launch list + one full capture of k_step.
# usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --inner 6 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
