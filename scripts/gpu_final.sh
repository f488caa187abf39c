#!/bin/bash
# Round-end style visit: smoke, all GPU tests, headline bench, reference arm, ncu evidence of the fast
# kernel and of the lean (reference-precision) kernel.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r02f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
echo "== bench c4"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err; cut -c1-600 gpurun_out/bench_${TAG}.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_${TAG}.json 2> gpurun_out/bench_reference_${TAG}.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_reference_${TAG}.json
echo "== ncu k_step_fast"; bash scripts/gpu_profile.sh $TAG > gpurun_out/profile_$TAG.log 2>&1; tail -3 gpurun_out/profile_$TAG.log
echo "== ncu fp64 lean"; NCU=1 bash scripts/gpu_fp64.sh $TAG
ls gpurun_out | wc -l
