#!/bin/bash
# Per-CTA timeline of k_step_fast from a -DSPGG_TRACE build: bash scripts/gpu_trace.sh <tag>   (build/libspgg_<tag>.so)
mkdir -p gpurun_out
SPGG_B200_LIB=$PWD/build/libspgg_$1.so SPGG_TRACE_FILE=$PWD/gpurun_out/trace_$1 timeout 200 python bench.py --steps 1 --warmup 1 --inner 10 --no-cpu-baseline > gpurun_out/trace_$1.log 2>&1
python scripts/trace_stats.py gpurun_out/trace_$1.* 2>&1 | grep -v "\.log" | tee gpurun_out/trace_$1_stats.txt
