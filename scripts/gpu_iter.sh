#!/bin/bash
# fast iteration: fast-path parity subset + bench + ncu profile
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "fast or L4096 or fp32_philox or chunking or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
bash scripts/gpu_profile.sh $TAG
