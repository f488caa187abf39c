#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench c4"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
for w in c1 c2 sweep; do SPGG_NO_RESIDENT=1 timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --inner 200 > gpurun_out/bench_${w}_periter.log 2> gpurun_out/bench_${w}_periter.err; done
SPGG_NO_RESIDENT=1 SPGG_NO_PDL=1 timeout 600 python bench.py --workload c1 --steps 3 --warmup 2 --inner 200 > gpurun_out/bench_c1_periter_nopdl.log 2>&1
