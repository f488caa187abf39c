#!/bin/bash
mkdir -p gpurun_out
for tag in static dyn; do
  export SPGG_B200_LIB=$PWD/build/lib_$tag.so
  SPGG_TRACE_FILE=$PWD/gpurun_out/trace_$tag timeout 300 python bench.py --steps 1 --warmup 1 --inner 10 --no-cpu-baseline > gpurun_out/trace_$tag.log 2>&1
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$tag.log 2> gpurun_out/var_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_$tag.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('$tag value %.3e k_step %.1f us frac %.3f gmax %.1f us whole %.3f'%(d['value'],r['kernel_us'],r['frac'],r['gmax_kernel_us'],r['whole_step_frac']))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/var_$tag.err').read()[-800:])
PY
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "fast or L4096 or fp32_philox or chunking or batched or strips_equal_single_lattice_gloo" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest $tag rc=$?"; tail -12 gpurun_out/pytest_$tag.log
done
